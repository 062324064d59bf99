#!/usr/bin/env python
"""Multi-GPU correctness check, run under torchrun with N >= 2 ranks (NCCL):

1. complex-sharded screening (`parallel.screen`) returns, on every rank, the
   same per-complex scores as one rank scoring everything itself;
2. data-parallel training: after `make_data_parallel`, the all-reduced
   gradient on every rank equals the single-process gradient of the
   concatenated batch (to 1e-5), and replicas stay bit-identical after a step.
"""
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import pointvs_b200 as pv  # noqa: E402
from pointvs_b200 import parallel  # noqa: E402
from pointvs_b200.synthetic import synthetic_complex  # noqa: E402


def main():
    rank, local_rank, world, dev = parallel.init_from_env()
    assert world >= 2, 'run under torchrun with at least 2 ranks'
    kw = dict(dim_input=13, dim_output=1, k=64, num_layers=4,
              edge_attention=True, node_attention=True, residual=True,
              normalize=True, tanh=True, graphnorm=False,
              model_task='classification')
    torch.manual_seed(0)
    model = pv.MultitaskSatorrasEGNN(Path('/tmp/pvs_multi'), 1e-3, 0, None,
                                     None, silent=True, **kw).to(dev)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('coord_mlp.2.weight'):
                p.mul_(1000.0)

    # ---- 1. sharded screening ----
    rng = np.random.default_rng(0)
    complexes = [synthetic_complex(500 + i, int(rng.integers(300, 700)), 20)
                 for i in range(4 * world + 3)]
    sharded = parallel.screen(model, complexes, batch_size=4)
    whole = []
    model.eval()
    with torch.no_grad():
        for c, b, f in complexes:
            batch = pv.PackedBatch.from_arrays(c, b, f, [0, len(c)], 4.0, 4.0,
                                               device=dev)
            whole.append(torch.sigmoid(model(batch)).reshape(-1).cpu().numpy())
    whole = np.stack(whole)
    err = float(np.max(np.abs(sharded - whole) / np.abs(whole)))
    assert err < 1e-5, f'sharded scores differ: {err}'

    # ---- 2. data-parallel gradient ----
    per_rank = 2
    all_c = [synthetic_complex(900 + i, 400, 20) for i in range(per_rank * world)]
    labels = torch.tensor([float(i % 2) for i in range(per_rank * world)],
                          device=dev)

    def batch_of(items, y):
        coords = np.concatenate([c for c, _, _ in items])
        bp = np.concatenate([b for _, b, _ in items])
        feats = np.concatenate([f for _, _, f in items])
        cptr = np.concatenate([[0], np.cumsum([len(c) for c, _, _ in items])])
        return pv.PackedBatch.from_arrays(coords, bp, feats, cptr, 4.0, 4.0,
                                          y=y, device=dev)

    model.train()
    # single-process reference gradient on the whole batch
    model.zero_grad()
    out = model(batch_of(all_c, labels)).reshape(-1)
    torch.nn.functional.binary_cross_entropy_with_logits(out, labels).backward()
    used = [p.grad is not None for p in model.parameters()]
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()
                     if p.grad is not None]).clone()
    # data-parallel: each rank its shard, then the flat all-reduce
    reducer = parallel.make_data_parallel(model)
    model.zero_grad(set_to_none=True)
    lo, hi = rank * per_rank, (rank + 1) * per_rank
    out = model(batch_of(all_c[lo:hi], labels[lo:hi])).reshape(-1)
    torch.nn.functional.binary_cross_entropy_with_logits(
        out, labels[lo:hi]).backward()
    reducer.sync()
    assert used == [p.grad is not None for p in model.parameters()]
    got = torch.cat([p.grad.reshape(-1) for p in model.parameters()
                     if p.grad is not None])
    gerr = float((got - ref).abs().max() / ref.abs().max())
    assert gerr < 1e-5, f'DP gradient differs from single-process: {gerr}'
    if rank == 0:
        print(f'multi-GPU check ok on {world} ranks: sharded-score rel err '
              f'{err:.2e}, DP-gradient rel err {gerr:.2e}', flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
