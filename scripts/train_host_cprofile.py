"""cProfile of the host side of the training step (where the Python time goes)."""
import cProfile, pstats, sys, io
sys.path.insert(0, '/root/repo')
import torch
from pathlib import Path
import pointvs_b200 as pv
from pointvs_b200.synthetic import synthetic_batch
from pointvs_b200 import parallel as _par
_rank, _lr, _world, _dev = _par.init_from_env()
kw = dict(dim_input=13, dim_output=1, k=64, num_layers=8, edge_attention=True,
          node_attention=True, residual=True, normalize=True, tanh=True,
          graphnorm=False, model_task='classification')
model = pv.MultitaskSatorrasEGNN(Path('/tmp/pvs_train'), 1e-3, 1e-4, None, None,
                                 silent=True, **kw).cuda().train()
model.set_math('bf16x3'); model.set_record_side_channels(False)
from pointvs_b200 import parallel
import os
if int(os.environ.get('WORLD_SIZE', '1')) > 1:
    parallel.make_data_parallel(model)
coords, bp, feats, cptr = synthetic_batch(0, 16, 1000, 30)
y = torch.tensor([i % 2 for i in range(16)], dtype=torch.float32, device='cuda')
c, b, f = (torch.from_numpy(a).cuda() for a in (coords, bp, feats))


def step():
    batch = pv.PackedBatch.from_arrays(c, b, f, cptr, 4.0, 4.0, y=y, device='cuda',
                                       edge_capacity='auto')
    batch.lig_fname = batch.rec_fname = [''] * 16
    yp, yt, _, _ = model.unpack_input_data_and_predict(batch)
    return model.backprop(yt, yp, sync=False)


for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
if _rank == 0:
    print(s.getvalue())
if _world > 1:
    torch.distributed.destroy_process_group()
