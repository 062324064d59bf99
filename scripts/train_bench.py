#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[3]): multitask EGNN, 8 layers
x 64 channels, fwd + bwd + clip + Adam, data-parallel over the visible ranks
with one flat NCCL gradient all-reduce per step.

    python scripts/train_bench.py [--batch 16] [--steps 10] [--math fp32]
    torchrun --nproc-per-node 2 scripts/train_bench.py ...

Prints one JSON line on rank 0.  Also asserts that every rank ends with
bit-identical parameters (the DP invariant)."""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import pointvs_b200 as pv  # noqa: E402
from pointvs_b200 import parallel  # noqa: E402
from pointvs_b200.synthetic import synthetic_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=16, help='complexes per GPU')
    ap.add_argument('--atoms', type=int, default=1000)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--math', default='fp32', choices=['fp32', 'bf16x3', 'bf16'])
    ap.add_argument('--exact-edges', action='store_true',
                    help='size the edge lists exactly (one host sync per step)')
    args = ap.parse_args()
    rank, local_rank, world, dev = parallel.init_from_env()
    kw = dict(dim_input=13, dim_output=1, k=64, num_layers=8,
              edge_attention=True, node_attention=True, residual=True,
              normalize=True, tanh=True, graphnorm=False,
              model_task='classification')
    torch.manual_seed(rank)          # replicas start different on purpose
    model = pv.MultitaskSatorrasEGNN(Path('/tmp/pvs_train'), 1e-3, 1e-4, None,
                                     None, silent=True, **kw).to(dev).train()
    model.set_math(args.math)
    model.set_record_side_channels(False)
    parallel.make_data_parallel(model)

    sets = []
    for s in range(3):
        coords, bp, feats, cptr = synthetic_batch(
            100_000 * rank + 1000 * s, args.batch, args.atoms, 30)
        y = torch.tensor([(i + s + rank) % 2 for i in range(args.batch)],
                         dtype=torch.float32, device=dev)
        sets.append((torch.from_numpy(coords).to(dev),
                     torch.from_numpy(bp).to(dev),
                     torch.from_numpy(feats).to(dev), cptr, y))

    overflow = []

    def step(i):
        coords, bp, feats, cptr, y = sets[i % 3]
        batch = pv.PackedBatch.from_arrays(coords, bp, feats, cptr, 4.0, 4.0,
                                           y=y, device=dev,
                                           edge_capacity=None if args.exact_edges else 'auto')
        batch.lig_fname = batch.rec_fname = [''] * args.batch
        y_pred, y_true, _, _ = model.unpack_input_data_and_predict(batch)
        overflow.append(batch.pvs_csr._overflow)
        return model.backprop(y_true, y_pred, sync=False)

    losses = [step(i) for i in range(args.warmup)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    ev0.record()
    t_host = time.perf_counter()
    for i in range(args.steps):
        losses.append(step(args.warmup + i))
    host_ms = (time.perf_counter() - t_host) * 1e3 / args.steps
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    assert int(torch.stack(overflow).sum().item()) == 0, 'edge capacity exceeded'
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    same = True
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        ok = torch.tensor([float(torch.equal(ref, flat))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(ok.item())
    if rank == 0:
        print(json.dumps({
            'metric': 'training complexes/s (multitask EGNN 8x64, fwd+bwd+clip+Adam)',
            'value': args.batch * world * args.steps / (float(ms) * 1e-3),
            'unit': 'complexes/s', 'n_gpus': world, 'steps': args.steps,
            'ms_per_step': float(ms) / args.steps, 'host_submit_ms_per_step': host_ms,
            'math': args.math,
            'batch_per_gpu': args.batch, 'first_loss': float(losses[0]),
            'last_loss': float(losses[-1]), 'replicas_identical': same,
            'allreduce_floats': sum(p.numel() for p in model.parameters())}),
            flush=True)
    assert same, 'replicas diverged'
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
