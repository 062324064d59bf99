"""The scheduling switches (launch chains, TMA stores, the side stream of the
stacked backward, the lean training pass) change HOW the kernels are issued,
not what they compute: scores, gradients and trained parameters are bitwise
equal with every switch on or off.  The switches are read once per process,
so every variant runs in its own interpreter."""
import hashlib
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]

SCRIPT = r'''
import hashlib, sys
sys.path.insert(0, %r)
import torch
from pathlib import Path
import pointvs_b200 as pv
from tests import gpu_helpers as gh

kw = dict(dim_input=13, dim_output=1, k=64, num_layers=3, edge_attention=True,
          node_attention=True, residual=True, normalize=True, tanh=True,
          graphnorm=False, model_task='classification')
torch.manual_seed(0)
m = pv.MultitaskSatorrasEGNN(Path('/tmp/pvs_test_switches'), 1e-3, 1e-4, None, None,
                             silent=True, **kw).cuda()
m.set_math('bf16x3')
m.set_record_side_channels(False)
m.record_embed_coords = False
m.eval()
g = gh.synthetic_graph(77, 6, 700, 20, ragged=True, edge_capacity='auto')
with torch.no_grad():
    scores = m(g).clone()
m.train()
for s_ in range(4):
    g = gh.synthetic_graph(100 + s_, 5, 500, 20, ragged=True, edge_capacity='auto')
    g.y = torch.tensor([float(i %% 2) for i in range(5)], device='cuda')
    g.lig_fname = g.rec_fname = [''] * 5
    yp, yt, _, _ = m.unpack_input_data_and_predict(g)
    m.backprop(yt, yp, sync=False)
torch.cuda.synchronize()
flat = torch.cat([scores.reshape(-1)] + [p.detach().reshape(-1) for p in m.parameters()])
assert torch.isfinite(flat).all()
print('HASH', hashlib.sha256(flat.cpu().numpy().tobytes()).hexdigest())
''' % str(ROOT)


def _run(extra_env):
    env = dict(os.environ)
    for k in ('PVS_NO_PDL', 'PVS_NO_TMA', 'PVS_NO_BWD_SIDE', 'PVS_STACK_TRAIN',
              'PVS_NODE_TC_TMA', 'PVS_EDGE_TC8'):
        env.pop(k, None)
    env.update(extra_env)
    out = subprocess.run([sys.executable, '-c', SCRIPT], env=env, cwd=str(ROOT),
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('HASH ')]
    assert len(lines) == 1, out.stdout[-2000:]
    return lines[0]


def test_scheduling_switches_do_not_change_a_single_bit():
    base = _run({})
    for env in ({'PVS_NO_PDL': '1'}, {'PVS_NO_TMA': '1'}, {'PVS_NO_BWD_SIDE': '1'},
                {'PVS_STACK_TRAIN': '0'}):
        assert _run(env) == base, env
