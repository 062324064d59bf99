"""One process per GPU: complex-sharded screening and data-parallel training.

The reference is single-process (no torch.distributed anywhere);
this is the multi-GPU layer the north star adds around the same model API.

* Screening / scoring: complexes are independent, so the list is split into
  one shard per rank and every rank scores its shard locally -- no collective
  on the data path.  Only the final per-complex scores are gathered (a few
  bytes per complex) so rank 0 can write the reference's predictions file.
* Training: replicas hold identical parameters.  Every parameter gradient
  lives in one flat fp32 arena (236 497 floats = 0.95 MB for the 8x64 model)
  that the backward kernels write directly; each EGNN layer's slice is
  averaged over NCCL/NVLink as soon as that layer's backward has been issued
  (overlapping the next layer's backward), the rest right after, all *before*
  `clip_grad_value_` and the optimiser step, so the clip sees the averaged
  gradient exactly as a single process with the whole batch would (reference
  order of operations: point_neural_network_base.py:417-429).  No pack / unpack
  copies: `p.grad` is a view of the arena.
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


class GradArena:
    """One flat fp32 buffer with a slot per parameter of `module`.

    The backward passes (pointvs_b200.backward) write their parameter
    gradients straight into the slots, so a training step needs one memset
    instead of a zero-filled tensor per layer, and a data-parallel step
    all-reduces slices of the buffer with no pack / unpack copies: after the
    reduction every `p.grad` is simply re-pointed at its slot.  Slots follow
    `module.parameters()` order, so the parameters of one EGNN layer are
    contiguous and a layer's gradients can be reduced as soon as its backward
    has been issued, while the next layer's backward runs (K5 in SURVEY.md).
    """

    ALIGN = 4          # floats: every slot starts on a 16-byte boundary

    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.slots, off = {}, 0
        for p in self.params:
            self.slots[p.data_ptr()] = (off, p.numel(), tuple(p.shape))
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.numel = off
        dev = self.params[0].device if self.params else torch.device('cpu')
        self.flat = torch.zeros(max(1, off), dtype=torch.float32, device=dev)
        self.group = None
        self.reduce_in_backward = False   # set by GradAllReducer
        self._handed, self._works, self._reduced = set(), [], []
        # one long-lived view per slot: `p.grad` of a parameter whose gradient
        # the arena-aware backward passes produce IS this view, step after step
        self.views = {p.data_ptr(): self.flat[o:o + n].view(shape)
                      for p, (o, n, shape) in ((p, self.slots[p.data_ptr()])
                                               for p in self.params)}
        self._granted, self._assigned = set(), set()
        self._pending = None
        self._ptrs = [p.data_ptr() for p in self.params]
        self._all_keys = set(self._ptrs)

    def matches(self, module, full=False):
        """Is this still the arena of `module`'s parameters?  The per-step
        check only looks at the parameters the arena already holds (storage
        moved by `.to()`, or frozen since): walking `module.parameters()` costs
        more host time than the rest of `begin_step`.  A parameter that was
        frozen when the arena was built and is trainable now simply gets its
        gradient through plain autograd.  full=True walks the module."""
        if full:
            ps = [p for p in module.parameters() if p.requires_grad]
            return len(ps) == len(self.params) and all(
                p.data_ptr() in self.slots and p.device == self.flat.device
                for p in ps)
        ptrs = self._ptrs
        for i, p in enumerate(self.params):
            if p.data_ptr() != ptrs[i] or not p.requires_grad:
                return False
        return True

    def clamp_(self, clip_value):
        """`clip_grad_value_` over every gradient that lives in the arena: one
        kernel over the flat buffer (slots without a gradient this step are
        zero).  Returns the parameters whose gradient is NOT an arena view, for
        the caller to clip the usual way."""
        self.flat.clamp_(-clip_value, clip_value)
        if self._assigned == self._all_keys:
            return []
        return [p for p in self.params
                if p.grad is not None and p.data_ptr() not in self._assigned]

    def begin_step(self):
        """Zero every slot (one kernel) and forget the previous step.  Gradients
        that plain autograd left on a parameter in the previous step are
        dropped here (what `optimiser.zero_grad()` would do); gradients that
        are views of the arena stay attached and are simply zeroed."""
        for p in self.params:
            if p.grad is not None and p.data_ptr() not in self._assigned:
                p.grad = None
        self.flat.zero_()
        self._handed.clear()
        self._granted.clear()
        self._works.clear()
        self._reduced.clear()
        self._pending = None

    def view(self, param):
        return self.views[param.data_ptr()]

    def grad_view(self, param):
        """`param`'s (zeroed) slot for a backward pass to accumulate into, or
        None when the parameter has no slot or has already handed its slot out
        in this step (a parameter used twice: autograd must add the second
        contribution itself)."""
        if param is None:
            return None
        key = param.data_ptr()
        if key not in self.slots or key in self._handed:
            return None
        self._handed.add(key)
        return self.views[key]

    def hand_out_all(self, keys):
        """Fast path of a backward pass with a cached plan: hand out the slots
        `keys` (a frozenset of data_ptrs) in one go.  False (and nothing
        changed) if any of them has already been handed out in this step."""
        if not self._handed.isdisjoint(keys):
            return False
        self._handed |= keys
        return True

    def grant_all(self, keys):
        self._granted |= keys

    def grant(self, param):
        """The backward pass that was handed `param`'s slot reports that the
        slot now holds a real gradient (parameters that cannot influence the
        loss are handed out but never granted: their `.grad` stays None, as
        under the reference's autograd)."""
        self._granted.add(param.data_ptr())

    def attach_grads(self):
        """After backward: `p.grad` = its arena view for every granted
        parameter, None for a parameter that was granted in an earlier step but
        not in this one (e.g. the other head after a task switch).  No tensor is
        created and no kernel runs; in steady state this is two set compares."""
        if self._granted == self._assigned and all(
                p.grad is not None for p in self.params
                if p.data_ptr() in self._assigned):
            return
        for p in self.params:
            key = p.data_ptr()
            if key in self._granted:
                if p.grad is None or p.grad.data_ptr() != self.views[key].data_ptr():
                    p.grad = self.views[key]
            elif key in self._assigned:
                p.grad = None
        self._assigned = set(self._granted)

    def span(self, params):
        """[lo, hi) of the slots of `params` (None entries are skipped), the
        alignment padding after the last slot included: the spans of
        consecutive layers are then adjacent and merge into one bucket, and no
        few-float gaps are left for `finish_reduce` to reduce one by one."""
        lo, hi = None, None
        for p in params:
            if p is None or p.data_ptr() not in self.slots:
                continue
            off, n, _ = self.slots[p.data_ptr()]
            end = off + (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            lo = off if lo is None else min(lo, off)
            hi = end if hi is None else max(hi, end)
        return lo, hi

    # -- data-parallel reduction -------------------------------------------------
    def _world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def _all_reduce(self, lo, hi):
        t = self.flat[lo:hi]
        if dist.get_backend(self.group) == 'nccl':
            self._works.append((dist.all_reduce(
                t, op=dist.ReduceOp.AVG, group=self.group, async_op=True), None))
        else:       # gloo (CPU tests) has no AVG
            self._works.append((dist.all_reduce(
                t, op=dist.ReduceOp.SUM, group=self.group, async_op=True), t))
        self._reduced.append((lo, hi))

    # Floats per all-reduce issued from inside backward.  A collective costs
    # ~0.1 ms of HOST time to issue and ~30 us on NVLink for the whole 0.95 MB
    # arena of the 8 x 64 model, so one call per layer (8 + 1 per step) made
    # the host the bottleneck of the data-parallel step; adjacent layer spans
    # are merged until a bucket is this large (default: a third of the arena).
    bucket_floats = None

    def reduce_async(self, lo, hi):
        """Average flat[lo:hi) over the ranks on the collective's own stream;
        called from inside backward right after the kernels that fill the
        range have been issued.  Adjacent ranges are merged into buckets."""
        if lo is None or hi is None or hi <= lo or self._world() == 1:
            return
        pend = self._pending
        if pend is not None and (pend[0] == hi or pend[1] == lo):
            pend = (min(pend[0], lo), max(pend[1], hi))
        else:
            if pend is not None:
                self._all_reduce(*pend)
            pend = (lo, hi)
        bucket = self.bucket_floats if self.bucket_floats is not None \
            else max(1, self.numel // 3)
        if pend[1] - pend[0] >= bucket:
            self._all_reduce(*pend)
            pend = None
        self._pending = pend

    def finish_reduce(self):
        """Reduce whatever reduce_async has not covered, then wait."""
        world = self._world()
        if world > 1:
            # an unfinished bucket is simply left uncovered: it merges with the
            # slots around it (embedding, heads) into one gap below
            self._pending = None
            covered = sorted(self._reduced)
            pos, gaps = 0, []
            for lo, hi in covered:
                if lo > pos:
                    gaps.append((pos, lo))
                pos = max(pos, hi)
            if pos < self.numel:
                gaps.append((pos, self.numel))
            for lo, hi in gaps:
                self._all_reduce(lo, hi)
            for work, t in self._works:
                work.wait()
                if t is not None:
                    t.div_(world)
        self._works.clear()
        self._reduced.clear()

    def in_reduced_span(self, param):
        off, n, _ = self.slots[param.data_ptr()]
        return any(lo < off + n and off < hi for lo, hi in self._reduced)


class GradAllReducer:
    """Averages the gradients of `module` across ranks through a GradArena.
    Install with `attach(model)`: the training loop's `backprop()` then calls
    `sync()` between backward and clip/step.

    Arena-aware backward passes (the EGNN model on CUDA) have already written
    into the arena and started per-layer reductions; gradients produced by
    plain autograd (any other module) are copied into their slots first.
    Either way every `p.grad` ends up as a view of the reduced arena, and a
    parameter whose gradient is None stays None (an optimiser with weight
    decay must skip it, exactly as a single process would)."""

    def __init__(self, module, group=None):
        self.module = module
        self.group = group
        self.arena = None
        self.numel = sum(p.numel() for p in module.parameters()
                         if p.requires_grad)

    def _arena(self):
        if self.arena is None or not self.arena.matches(self.module):
            arena = getattr(self.module, '_grad_arena', None)
            if arena is None or not arena.matches(self.module):
                arena = GradArena(self.module)
                if hasattr(self.module, '_grad_arena'):
                    self.module._grad_arena = arena
            arena.group = self.group
            arena.reduce_in_backward = True
            self.arena = arena
        return self.arena

    def sync(self):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        arena = self._arena()
        with torch.no_grad():
            for p in arena.params:
                # a slot handed to an arena-aware backward already holds this
                # step's gradient (and may be mid-reduction): never write it
                if p.grad is None or p.data_ptr() in arena._handed:
                    continue
                if p.data_ptr() in arena._assigned:
                    continue
                v = arena.view(p)
                if p.grad.data_ptr() != v.data_ptr():
                    if arena.in_reduced_span(p):
                        raise RuntimeError(
                            'a plain-autograd gradient lies inside a span that '
                            'was reduced during backward')
                    v.copy_(p.grad)
            arena.finish_reduce()
            for p in arena.params:
                if p.grad is not None:
                    p.grad = arena.view(p)

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
