"""Kernel times of one scoring pass of the bench workload (torch profiler,
in situ: warm L2, kernels back to back)."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from pathlib import Path
import bench
import pointvs_b200 as pv
from pointvs_b200.synthetic import synthetic_batch
from torch.profiler import profile, ProfilerActivity

math = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
torch.manual_seed(0)
model = pv.SartorrasEGNN(Path('/tmp/pvs_bench'), 0, 0, None, None, silent=True,
                         **bench.MODEL_KW).cuda().eval()
model.set_math(math)
model.set_record_side_channels(False)
model.record_embed_coords = False
coords, bp, feats, cptr = synthetic_batch(0, 128, 1000, 30)
c, b, f = (torch.from_numpy(a).cuda() for a in (coords, bp, feats))


def step():
    batch = pv.PackedBatch.from_arrays(c, b, f, cptr, bench.EDGE_RADIUS, bench.EDGE_RADIUS,
                                       device='cuda', edge_capacity='auto')
    with torch.no_grad():
        return model(batch)


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=20, max_name_column_width=60))
