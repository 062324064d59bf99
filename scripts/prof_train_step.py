import sys, collections
sys.path.insert(0, '/root/repo')
import torch, traceback
from pathlib import Path
import pointvs_b200 as pv
from pointvs_b200.synthetic import synthetic_batch
kw = dict(dim_input=13, dim_output=1, k=64, num_layers=8, edge_attention=True, node_attention=True, residual=True, normalize=True, tanh=True, graphnorm=False, model_task='classification')
dev='cuda'
model = pv.MultitaskSatorrasEGNN(Path('/tmp/pvs_train'), 1e-3, 1e-4, None, None, silent=True, **kw).to(dev).train()
model.set_math('bf16x3'); model.set_record_side_channels(False)
coords, bp, feats, cptr = synthetic_batch(0, 16, 1000, 30)
y = torch.tensor([i % 2 for i in range(16)], dtype=torch.float32, device=dev)
c,b,f = torch.from_numpy(coords).to(dev), torch.from_numpy(bp).to(dev), torch.from_numpy(feats).to(dev)
def step():
    batch = pv.PackedBatch.from_arrays(c, b, f, cptr, 4.0, 4.0, y=y, device=dev, edge_capacity='auto')
    batch.lig_fname = batch.rec_fname = [''] * 16
    yp, yt, _, _ = model.unpack_input_data_and_predict(batch)
    return model.backprop(yt, yp, sync=False)
for _ in range(3): step()
counts = collections.Counter()
orig_zeros, orig_zeros_like, orig_zero_ = torch.zeros, torch.zeros_like, torch.Tensor.zero_
def wrap(fn, name):
    def inner(*a, **k):
        st = traceback.extract_stack(limit=4)
        counts[(name, tuple((s.filename.split('/')[-1], s.lineno) for s in st[:-1]))] += 1
        return fn(*a, **k)
    return inner
torch.zeros = wrap(orig_zeros, 'zeros'); torch.zeros_like = wrap(orig_zeros_like, 'zeros_like'); torch.Tensor.zero_ = wrap(orig_zero_, 'zero_')
torch.full = wrap(torch.full, 'full'); torch.ones = wrap(torch.ones, 'ones')
step()
torch.cuda.synchronize()
for k_, v in counts.most_common(15): print(v, k_)
from torch.profiler import profile, ProfilerActivity
torch.zeros, torch.zeros_like, torch.Tensor.zero_ = orig_zeros, orig_zeros_like, orig_zero_
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=45, max_name_column_width=70))
