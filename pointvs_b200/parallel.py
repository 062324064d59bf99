"""One process per GPU: complex-sharded screening and data-parallel training.

The reference is single-process (no torch.distributed anywhere);
this is the multi-GPU layer the north star adds around the same model API.

* Screening / scoring: complexes are independent, so the list is split into
  one shard per rank and every rank scores its shard locally -- no collective
  on the data path.  Only the final per-complex scores are gathered (a few
  bytes per complex) so rank 0 can write the reference's predictions file.
* Training: replicas hold identical parameters; after backward the gradients
  are flattened into one fp32 buffer (236 497 floats = 0.95 MB for the 8x64
  model), summed with ONE all-reduce over NCCL/NVLink and divided by the world
  size, *before* `clip_grad_value_` and the optimiser step, so the clip sees
  the averaged gradient exactly as a single process with the whole batch would
  (reference order of operations: point_neural_network_base.py:417-429).
  The message is latency-bound on NVSwitch, so it is a single bucket.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous block [lo, hi) of `n_items` owned by `rank`."""
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_size(sizes, world_size):
    """Greedy longest-first partition of ragged work items (e.g. atoms per
    complex) into `world_size` shards of near-equal total size.  Returns a
    list of index arrays, each sorted ascending so per-rank order is stable."""
    sizes = np.asarray(sizes, dtype=np.int64)
    order = np.argsort(-sizes, kind='stable')
    loads = np.zeros(world_size, dtype=np.int64)
    shards = [[] for _ in range(world_size)]
    for idx in order:
        r = int(np.argmin(loads))
        shards[r].append(int(idx))
        loads[r] += sizes[idx]
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


def gather_scores(local_indices, local_scores, n_total, group=None):
    """All ranks contribute (index, score) pairs; every rank gets the dense
    [n_total, ...] score array back in the original order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    local_scores = np.asarray(local_scores)
    if world == 1:
        out = np.zeros((n_total,) + local_scores.shape[1:], local_scores.dtype)
        out[np.asarray(local_indices)] = local_scores
        return out
    payload = (np.asarray(local_indices), local_scores)
    gathered = [None] * world
    dist.all_gather_object(gathered, payload, group=group)
    out = np.zeros((n_total,) + local_scores.shape[1:], local_scores.dtype)
    for idx, sc in gathered:
        if len(idx):
            out[idx] = sc
    return out


class GradAllReducer:
    """Averages the gradients of `module` across ranks with one flat
    all-reduce.  Install with `attach(model)`: the training loop's
    `backprop()` then calls it between backward and clip/step."""

    def __init__(self, module, group=None):
        self.module = module
        self.group = group
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self._flat = None

    def _buffer(self, device):
        if self._flat is None or self._flat.device != device:
            self._flat = torch.zeros(self.numel, dtype=torch.float32,
                                     device=device)
        return self._flat

    def sync(self):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        device = self.params[0].device
        flat = self._buffer(device)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                flat[off:off + n].zero_()
            else:
                flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(dist.get_world_size(self.group))
        off = 0
        for p in self.params:
            n = p.numel()
            # a parameter unused on this rank is unused on every rank (same
            # model, same task): leave its grad None so the optimiser skips
            # it exactly as a single process would
            if p.grad is not None:
                p.grad.copy_(flat[off:off + n].view_as(p))
            off += n

    def attach(self, model=None):
        (model or self.module).sync_gradients = self.sync
        return self


def broadcast_parameters(module, src=0, group=None):
    """Make every replica start from rank `src`'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def make_data_parallel(model, group=None):
    """Replicate `model` (a PointNeuralNetworkBase) for data-parallel
    training: identical start, averaged gradients every step."""
    broadcast_parameters(model, 0, group)
    return GradAllReducer(model, group).attach(model)


def init_from_env(backend=None):
    """torchrun-style initialisation (RANK / LOCAL_RANK / WORLD_SIZE /
    MASTER_*).  Returns (rank, local_rank, world_size, device)."""
    import os
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
        device = torch.device('cuda', local_rank)
    else:
        device = torch.device('cpu')
    if world > 1 and not dist.is_initialized():
        backend = backend or ('nccl' if device.type == 'cuda' else 'gloo')
        kwargs = {'device_id': device} if backend == 'nccl' else {}
        dist.init_process_group(backend, **kwargs)
    return rank, local_rank, world, device


def screen(model, complexes, batch_size=128, inter_radius=4.0,
           intra_radius=4.0, group=None, activation='sigmoid'):
    """Score a list of complexes [(coords f64 [N,3], bp [N], feats [N,F]), ...]
    sharded by complex across the ranks of `group`.  Returns the dense score
    array (same on every rank)."""
    from .pipeline import ScoreStream
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    sizes = [len(c[0]) for c in complexes]
    mine = shard_by_size(sizes, world)[rank]
    model.eval()
    # exact edge lists (one 4-byte read-back per batch): callers' complexes
    # may be denser than the capacity heuristic assumes
    stream = ScoreStream(model, inter_radius, intra_radius, depth=3,
                         edge_capacity=None, activation=activation)
    for lo in range(0, len(mine), batch_size):
        idx = mine[lo:lo + batch_size]
        coords = np.concatenate([complexes[i][0] for i in idx])
        bp = np.concatenate([complexes[i][1] for i in idx])
        feats = np.concatenate([complexes[i][2] for i in idx])
        cptr = np.concatenate([[0], np.cumsum([sizes[i] for i in idx])])
        stream.submit(coords, bp, feats, cptr)
    scores = [sc for _, sc in stream.drain()]
    local = np.concatenate(scores) if scores else \
        np.zeros((0, 1), dtype=np.float32)
    return gather_scores(mine, local, len(complexes), group)
