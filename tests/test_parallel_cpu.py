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


def _tiny_module():
    import torch
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3),
                               torch.nn.Linear(3, 2))


def test_arena_spans_are_adjacent_and_merge_into_buckets(monkeypatch):
    """Layer spans include their alignment padding, so consecutive spans touch
    and `reduce_async` merges them until a bucket is full; `finish_reduce`
    covers the rest in ONE call (no few-float gaps reduced one by one)."""
    from pointvs_b200.parallel import GradArena
    m = _tiny_module()
    arena = GradArena(m)
    spans = [arena.span(list(layer.parameters())) for layer in m]
    for (lo0, hi0), (lo1, _) in zip(spans, spans[1:]):
        assert hi0 == lo1                      # padded end == next slot's offset
    assert spans[0][0] == 0 and spans[-1][1] == arena.numel
    calls = []
    monkeypatch.setattr(arena, '_world', lambda: 2)
    monkeypatch.setattr(arena, '_all_reduce',
                        lambda lo, hi: (calls.append((lo, hi)),
                                        arena._reduced.append((lo, hi))))
    arena.bucket_floats = spans[2][1] - spans[1][0]      # layers 2 + 1 fill a bucket
    for lo, hi in reversed(spans):                       # backward order
        arena.reduce_async(lo, hi)
    # layers 2 and 1 merged into one call; layer 0 alone is a bucket
    assert calls == [(spans[1][0], spans[2][1]), spans[0]]
    arena.finish_reduce()
    assert len(calls) == 2                               # nothing left over
    # a bucket larger than the arena: nothing is issued from inside backward and
    # the unfinished bucket goes out as ONE call at the end
    del calls[:]
    arena.bucket_floats = 10 * arena.numel
    for lo, hi in reversed(spans):
        arena.reduce_async(lo, hi)
    assert calls == []
    arena.finish_reduce()
    assert calls == [(0, arena.numel)]


def test_arena_clamp_equals_clip_grad_value():
    import torch
    from pointvs_b200.parallel import GradArena
    m = _tiny_module()
    arena = GradArena(m)
    arena.begin_step()
    g = torch.Generator().manual_seed(1)
    for p in m.parameters():
        v = arena.grad_view(p)
        v.copy_(torch.randn(p.shape, generator=g) * 3)
        arena.grant(p)
    arena.attach_grads()
    want = [p.grad.clone().clamp_(-1.0, 1.0) for p in m.parameters()]
    assert arena.clamp_(1.0) == []
    for p, w in zip(m.parameters(), want):
        assert p.grad.data_ptr() == arena.view(p).data_ptr()
        assert torch.equal(p.grad, w)
    assert arena.matches(m) and arena.matches(m, full=True)
    m[0].weight.requires_grad_(False)
    assert not arena.matches(m)


def test_auto_edge_capacity_grows_with_the_radius():
    from pointvs_b200.graph import auto_edge_capacity
    assert auto_edge_capacity(1000, 4.0, 4.0) == 24 * 1000
    assert auto_edge_capacity(1000, 4.0, 2.0) == 24 * 1000
    assert auto_edge_capacity(10, 6.0, 4.0) == 81 * 10
