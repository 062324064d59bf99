#!/usr/bin/env python
"""Diagnostic: which part of bench.py's instrumentation perturbs the timed loop."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
import pointvs_b200 as pv
from pointvs_b200 import egnn as egnn_mod
from pointvs_b200.synthetic import synthetic_batch

dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = pv.SartorrasEGNN(Path('/tmp/pvs_bench'), 0, 0, None, None, silent=True,
                         **bench.MODEL_KW).to(dev).eval()
model.set_math('bf16x3'); model.set_record_side_channels(False); model.record_embed_coords = False
sets = []
for s in range(3):
    c, b, f, p = synthetic_batch(10_000 * s, 128, 1000, 30)
    sets.append((torch.from_numpy(c).to(dev), torch.from_numpy(b).to(dev), torch.from_numpy(f).to(dev), p))

def step(i):
    c, b, f, p = sets[i % 3]
    batch = pv.PackedBatch.from_arrays(c, b, f, p, 4.0, 4.0, device=dev, edge_capacity='auto')
    with torch.no_grad():
        return model(batch), batch.pvs_csr

def run(name, steps=20, timer=False, sampler=False, throttle=False, accum=False):
    for i in range(5): step(i)
    torch.cuda.synchronize()
    smp = bench.ClockSampler(0) if sampler else None
    if smp and smp.ok: smp.start()
    if timer: egnn_mod.STAGE_TIMER = bench.EdgeKernelTimer(torch, steps * 8)
    ed = torch.zeros(1, dtype=torch.int64, device=dev); ov = torch.zeros(1, dtype=torch.int32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fl = []
    t0 = time.perf_counter(); e0.record()
    for i in range(steps):
        if throttle and len(fl) >= 3: fl.pop(0).synchronize()
        _, csr = step(i)
        if accum: ed += csr.n_edges_dev; ov += csr._overflow
        del csr
        if throttle:
            d = torch.cuda.Event(); d.record(); fl.append(d)
    e1.record(); host = (time.perf_counter() - t0) * 1e3 / steps
    torch.cuda.synchronize()
    egnn_mod.STAGE_TIMER = None
    if smp: smp.stop()
    print(f'{name:28s} {e0.elapsed_time(e1) / steps:7.3f} ms/step  host {host:6.3f}', flush=True)

run('bare'); run('bare again'); run('timer', timer=True); run('sampler', sampler=True)
run('throttle', throttle=True); run('accum', accum=True)
run('all', timer=True, sampler=True, throttle=True, accum=True)
run('timer+accum', timer=True, accum=True)
run('bare 100 steps', steps=100)
