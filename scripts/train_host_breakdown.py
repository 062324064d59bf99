#!/usr/bin/env python
"""Host-side time of one training step by section (no device syncs inside the
step): tells whether the step is paced by the host or by the device."""
import os, sys, time, collections, math
sys.path.insert(0, '/root/repo')
import torch
from pathlib import Path
import pointvs_b200 as pv
from pointvs_b200 import backward as _bw
from pointvs_b200.synthetic import synthetic_batch

kw = dict(dim_input=13, dim_output=1, k=64, num_layers=8, edge_attention=True,
          node_attention=True, residual=True, normalize=True, tanh=True,
          graphnorm=False, model_task='classification')
dev = 'cuda'
model = pv.MultitaskSatorrasEGNN(Path('/tmp/pvs_train'), 1e-3, 1e-4, None, None,
                                 silent=True, **kw).to(dev).train()
model.set_math('bf16x3'); model.set_record_side_channels(False)
model._lean_training = os.environ.get('LEAN', '1') == '1'
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
coords, bp, feats, cptr = synthetic_batch(0, B, 1000, 30)
y = torch.tensor([i % 2 for i in range(B)], dtype=torch.float32, device=dev)
c, b, f = (torch.from_numpy(a).to(dev) for a in (coords, bp, feats))
T = collections.OrderedDict()


def tick(name, t0):
    t1 = time.perf_counter()
    T[name] = T.get(name, 0.0) + (t1 - t0)
    return t1


def step():
    t = time.perf_counter()
    batch = pv.PackedBatch.from_arrays(c, b, f, cptr, 4.0, 4.0, y=y, device=dev, edge_capacity='auto')
    batch.lig_fname = batch.rec_fname = [''] * B
    t = tick('graph build (from_arrays)', t)
    yp, yt, _, _ = model.unpack_input_data_and_predict(batch)
    t = tick('forward', t)
    loss = model.get_loss(yt, yp)
    t = tick('loss', t)
    arena = model._arena_for_step(loss)
    arena.begin_step()
    t = tick('arena begin', t)
    with _bw.use_arena(arena):
        loss.backward()
    t = tick('backward', t)
    arena.attach_grads()
    model.sync_gradients()
    t = tick('attach + sync_gradients', t)
    torch.nn.utils.clip_grad_value_(model.parameters(), 1.0)
    t = tick('clip', t)
    model.optimiser.step()
    t = tick('adam', t)


for _ in range(5):
    step()
for mode in ('free-running', 'sync before each step'):
    T.clear()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    w0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        if mode != 'free-running':
            torch.cuda.synchronize()
        step()
    host = time.perf_counter() - w0
    e1.record(); torch.cuda.synchronize()
    print(f'== {mode}: host {host / n * 1e3:.3f} ms/step, device {e0.elapsed_time(e1) / n:.3f} ms/step')
    for k_, v in T.items():
        print(f'   {k_:32s} {v / n * 1e3:7.3f} ms')
