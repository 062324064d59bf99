import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pathlib import Path
import bench
import pointvs_b200 as pv
from pointvs_b200 import data
from pointvs_b200.graph import radius_graph_batch
from pointvs_b200.synthetic import N_TYPES, synthetic_ligand_poses, synthetic_pocket
dev = torch.device('cuda')
torch.manual_seed(0)
model = pv.SartorrasEGNN(Path('/tmp/pvs_bench'), 0, 0, None, None, silent=True, **bench.MODEL_KW).cuda().eval()
model.set_math('bf16x3'); model.set_record_side_channels(False); model.record_embed_coords = False
n_lig, n_pocket, batch = 30, 800, 128
pxyz, ptypes, _ = synthetic_pocket(n_pocket, n_lig)
pocket = (torch.from_numpy(pxyz.copy()).to(dev), torch.ones(n_pocket, dtype=torch.uint8, device=dev),
          torch.from_numpy((ptypes + N_TYPES).astype(np.int16)).to(dev))
lig_xyz, lig_types = synthetic_ligand_poses(0, 128 * 64, n_lig)
emit = np.ones(n_lig, dtype=np.uint8)
zeros = np.zeros(batch, dtype=np.int32)
T = {}
def tick(k, t0):
    t1 = time.perf_counter(); T[k] = T.get(k, 0) + t1 - t0; return t1
def step(i):
    t = time.perf_counter()
    a = i * batch
    ligs = [data.Ligand(lig_xyz[p], emit, lig_types[p]) for p in range(a, a + batch)]
    t = tick('ligand objects', t)
    c, bp, f, cp = data.crop_batch(ligs, [pocket], zeros, 1e9, N_TYPES, True, dev)
    t = tick('crop_batch (K0)', t)
    csr = radius_graph_batch(c, bp, cp, 4.0, 4.0, device=dev, edge_capacity='auto')
    t = tick('radius_graph (K1)', t)
    pb = pv.PackedBatch(f, c.float(), csr, csr.complex_ptr)
    with torch.no_grad():
        out = model(pb)
    t = tick('model', t)
    return out
for i in range(8): step(i)
torch.cuda.synchronize(); T.clear()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
w0 = time.perf_counter(); e0.record()
n = 48
for i in range(n): step(i % 64)
host = time.perf_counter() - w0
e1.record(); torch.cuda.synchronize()
print(f'host {host/n*1e3:.3f} ms/step, device {e0.elapsed_time(e1)/n:.3f} ms/step')
for k, v in T.items(): print(f'  {k:22s} {v/n*1e3:.3f} ms')
