import sys, os, subprocess, json
sys.path.insert(0, '/root/repo')
import numpy as np
if len(sys.argv) > 1:
    import torch
    from pointvs_b200.graph import radius_graph_batch
    from pointvs_b200.synthetic import synthetic_batch
    coords, bp, _, cptr = synthetic_batch(0, 4)
    g = radius_graph_batch(coords, bp, cptr, 4.0, 4.0)
    np.savez(sys.argv[1], rp=g.row_ptr.cpu().numpy(), col=g.col.cpu().numpy(), attr=g.attr.cpu().numpy())
    sys.exit(0)
for name, lib in (('new', 'pointvs_b200/_C/libpvs_b200.so'), ('old', 'pointvs_b200/_C/variants/libpvs_r01.so')):
    subprocess.run([sys.executable, __file__, f'/tmp/k1_{name}.npz'], env=dict(os.environ, PVS_B200_LIB=os.path.abspath(lib)), check=True)
a, b = np.load('/tmp/k1_new.npz'), np.load('/tmp/k1_old.npz')
print('rp equal', np.array_equal(a['rp'], b['rp']), 'E', a['rp'][-1], b['rp'][-1])
da, db = np.diff(a['rp']), np.diff(b['rp'])
bad = np.nonzero(da != db)[0]
print('nodes with different degree', len(bad), bad[:20])
from pointvs_b200.synthetic import synthetic_batch
coords, bp, _, cptr = synthetic_batch(0, 4)
for i in bad[:8]:
    na = set(zip(a['col'][a['rp'][i]:a['rp'][i+1]].tolist(), a['attr'][a['rp'][i]:a['rp'][i+1]].tolist()))
    nb = set(zip(b['col'][b['rp'][i]:b['rp'][i+1]].tolist(), b['attr'][b['rp'][i]:b['rp'][i+1]].tolist()))
    miss = sorted(nb - na)
    print('node', i, 'deg new/old', da[i], db[i], 'missing', miss, 'extra', sorted(na - nb), 'xyz', coords[i], [ (j, coords[j], float(np.linalg.norm(coords[i]-coords[j]))) for j,_ in miss[:2]])
