"""Host-fed scoring without a per-batch stall.

`ScoreStream` is the public entry point for scoring batches that start in host
memory (the screening loop of inference.py:76-146 in the reference, where every
batch is `.to(device)`-ed, scored and its predictions pulled back with
`to_numpy` before the next one is touched).  Here every batch still crosses
the bus both ways, but:

  * inputs go host -> device on a copy stream into one of `depth` reusable
    staging slots, overlapping the previous batch's kernels;
  * the graph (K1) is built with a capacity-bounded edge list, so nothing is
    read back between the passes;
  * scores come back through pinned buffers with `non_blocking` copies and are
    only waited for when their slot is reused or at `drain()`.

Nothing is skipped: results are bit-identical to scoring the batches one at a
time (tests/test_gpu_pipeline.py).
"""
import numpy as np
import torch

from .graph import PackedBatch


class _Slot:
    def __init__(self):
        self.dev = {}          # name -> device staging tensor (grow-only)
        self.pin = {}          # name -> pinned host staging tensor
        self.out = None        # pinned scores
        self.h2d = None
        self.done = None
        self.meta = None       # (tag, shape of the scores)


class ScoreStream:
    """Pipelined scoring of host-resident packed batches.

    submit(coords f64 [N,3], bp [N], feats f32 [N,F], complex_ptr [B+1], tag)
    queues one batch (host arrays / tensors, or tensors already on the
    device); results() / drain() give (tag, scores ndarray [B, out]) in
    submission order.  `edges_scored` (device int64) counts the edges of
    everything submitted."""

    def __init__(self, model, inter_radius=4.0, intra_radius=4.0, depth=3,
                 edge_capacity='auto', activation=None):
        self.model = model
        self.device = next(model.parameters()).device
        if self.device.type != 'cuda':
            from ._cabi import PvsError
            raise PvsError('ScoreStream needs the model on a CUDA device '
                           '(no CPU fallback)')
        self.inter_radius, self.intra_radius = inter_radius, intra_radius
        self.edge_capacity = edge_capacity
        self.activation = activation
        self.copy_stream = torch.cuda.Stream(self.device)
        self.slots = [_Slot() for _ in range(max(1, depth))]
        self.count = 0
        self._pending = []     # slots in flight, oldest first
        self._ready = []       # finished (tag, scores)
        self._overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.edges_scored = torch.zeros(1, dtype=torch.int64,
                                        device=self.device)

    # -- staging ---------------------------------------------------------------
    @staticmethod
    def _as_cpu_tensor(a, dtype):
        t = torch.as_tensor(a)
        if t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous()

    def _stage(self, slot, name, src, dtype):
        """src (host) -> slot's device buffer on the copy stream.  A tensor
        that already lives on the device is used where it is."""
        if isinstance(src, torch.Tensor) and src.is_cuda:
            if src.dtype != dtype:
                # a conversion here would run on the copy stream, unordered
                # with the stream that produced `src`: make the caller do it
                raise TypeError(
                    f'ScoreStream: device-resident {name} must already be '
                    f'{dtype} (got {src.dtype}); convert it on your stream')
            return src
        src = self._as_cpu_tensor(src, dtype)
        n = src.shape[0]
        dev = slot.dev.get(name)
        if dev is None or dev.shape[0] < n or dev.shape[1:] != src.shape[1:]:
            cap = max(n, 1) if dev is None else max(n, int(1.25 * dev.shape[0]))
            dev = torch.empty((cap,) + tuple(src.shape[1:]), dtype=dtype,
                              device=self.device)
            slot.dev[name] = dev
        if not src.is_pinned():
            pin = slot.pin.get(name)
            if pin is None or pin.shape[0] < n or pin.shape[1:] != src.shape[1:]:
                pin = torch.empty((dev.shape[0],) + tuple(src.shape[1:]),
                                  dtype=dtype).pin_memory()
                slot.pin[name] = pin
            pin[:n].copy_(src)
            src = pin[:n]
        dev[:n].copy_(src, non_blocking=True)
        return dev[:n]

    # -- pipeline --------------------------------------------------------------
    def submit(self, coords, bp, feats, complex_ptr, tag=None):
        slot = self.slots[self.count % len(self.slots)]
        self.count += 1
        if slot.meta is not None:
            self._finish(slot)
        compute = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            # the slot's staging buffers were last read by the kernels of the
            # batch that used it before
            if slot.done is not None:
                self.copy_stream.wait_event(slot.done)
            coords_d = self._stage(slot, 'coords', coords, torch.float64)
            bp_d = self._stage(slot, 'bp', bp, torch.int32)
            feats_d = self._stage(slot, 'feats', feats, torch.float32)
            slot.h2d = torch.cuda.Event()
            slot.h2d.record(self.copy_stream)
        compute.wait_event(slot.h2d)
        batch = PackedBatch.from_arrays(
            coords_d, bp_d, feats_d, np.asarray(complex_ptr),
            self.inter_radius, self.intra_radius, device=self.device,
            edge_capacity=self.edge_capacity)
        with torch.no_grad():
            scores = self.model(batch)
        n_graphs = batch.num_graphs
        scores = scores.reshape(n_graphs, -1)
        if self.activation == 'sigmoid':
            scores = torch.sigmoid(scores)
        if self.edge_capacity is not None:
            self._overflow += batch.pvs_csr._overflow
        self.edges_scored += batch.pvs_csr.n_edges_dev
        if slot.out is None or slot.out.shape[0] < n_graphs or \
                slot.out.shape[1] != scores.shape[1]:
            slot.out = torch.empty((max(n_graphs, 1), scores.shape[1]),
                                   dtype=torch.float32).pin_memory()
        slot.out[:n_graphs].copy_(scores, non_blocking=True)
        slot.done = torch.cuda.Event()
        slot.done.record(compute)
        slot.meta = (tag, n_graphs)
        self._pending.append(slot)

    def _finish(self, slot):
        """Wait for the slot's scores and move them to the ready list (in
        submission order: older slots first)."""
        while self._pending:
            s = self._pending.pop(0)
            s.done.synchronize()
            tag, n_graphs = s.meta
            self._ready.append((tag, s.out[:n_graphs].numpy().copy()))
            s.meta = None
            if s is slot:
                break

    def results(self):
        """Scores finished so far (non-blocking), oldest first."""
        while self._pending and self._pending[0].done.query():
            self._finish(self._pending[0])
        out, self._ready = self._ready, []
        return out

    def drain(self):
        """Wait for everything submitted; -> [(tag, scores [B, out]), ...]."""
        if self._pending:
            self._finish(self._pending[-1])
        if int(self._overflow.item()):
            raise RuntimeError(
                'edge capacity exceeded: pass a larger edge_capacity (or None '
                'to size the edge list exactly)')
        out, self._ready = self._ready, []
        return out
