"""world_size-2 gloo tests (CPU) of the multi-process host logic: sharding,
score gathering, and the flat-gradient all-reduce used for data-parallel
training.  The CUDA kernels are not involved; the model here is a stand-in
nn.Module with the same `sync_gradients` / `backprop` contract."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from pointvs_b200 import parallel


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 128, 1001):
        for world in (1, 2, 3, 8):
            blocks = [parallel.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_size_balances_ragged_complexes():
    rng = np.random.default_rng(0)
    sizes = rng.integers(800, 1201, size=257)
    shards = parallel.shard_by_size(sizes, 8)
    all_idx = np.sort(np.concatenate(shards))
    np.testing.assert_array_equal(all_idx, np.arange(257))
    loads = np.array([sizes[s].sum() for s in shards])
    assert loads.max() - loads.min() <= 1200
    assert all(np.all(np.diff(s) > 0) for s in shards)


class _Toy(nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.net = nn.Sequential(nn.Linear(5, 7), nn.SiLU(), nn.Linear(7, 1))
        self.unused = nn.Parameter(torch.zeros(3))   # grad stays None

    def forward(self, x):
        return self.net(x).reshape(-1)

    def sync_gradients(self):
        pass


def _worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)
        model = _Toy()
        with torch.no_grad():       # diverge the replicas on purpose
            for p in model.parameters():
                p.add_(rank * 0.5)
        reducer = parallel.make_data_parallel(model)
        ref = _Toy()
        for p, q in zip(model.parameters(), ref.parameters()):
            assert torch.equal(p, q), 'broadcast_parameters failed'

        gen = torch.Generator().manual_seed(5)
        x = torch.randn(8, 5, generator=gen)
        y = (torch.rand(8, generator=gen) > 0.5).float()
        lo, hi = parallel.shard_range(8, rank, world)
        loss = nn.functional.binary_cross_entropy_with_logits(
            model(x[lo:hi]), y[lo:hi])
        loss.backward()
        model.sync_gradients()
        # single-process gradient of the whole batch (mean over 8 == mean of
        # the two per-rank means because the shards are equal-sized)
        loss_ref = nn.functional.binary_cross_entropy_with_logits(ref(x), y)
        loss_ref.backward()
        for (n, p), q in zip(model.named_parameters(), ref.parameters()):
            if q.grad is None:
                assert p.grad is None
            else:
                assert torch.allclose(p.grad, q.grad, atol=1e-6), n

        # score gathering for sharded screening
        sizes = np.array([5, 9, 2, 7, 7, 3, 8])
        mine = parallel.shard_by_size(sizes, world)[rank]
        local = np.stack([np.array([i * 10.0]) for i in mine]) if len(mine) \
            else np.zeros((0, 1))
        dense = parallel.gather_scores(mine, local, len(sizes))
        np.testing.assert_array_equal(dense[:, 0], np.arange(7) * 10.0)
        with open(os.path.join(tmpdir, f'ok{rank}'), 'w') as f:
            f.write('ok')
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / 'ok0').exists() and (tmp_path / 'ok1').exists()
