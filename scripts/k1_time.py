import sys, torch, time
sys.path.insert(0,'/root/repo')
from pointvs_b200.graph import radius_graph_batch
from pointvs_b200.synthetic import synthetic_batch
coords, bp, feats, cptr = synthetic_batch(0, 128, 1000, 30)
c = torch.from_numpy(coords).cuda(); b = torch.from_numpy(bp).cuda()
for _ in range(5): g = radius_graph_batch(c, b, cptr, 4.0, 4.0, edge_capacity='auto')
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(50): g = radius_graph_batch(c, b, cptr, 4.0, 4.0, edge_capacity='auto')
e1.record(); torch.cuda.synchronize()
print('K1 ms per batch', e0.elapsed_time(e1)/50, 'edges', int(g.n_edges_dev.item()), 'checksum', int(g.col[:int(g.n_edges_dev.item())].long().sum().item()))
