"""Radius graphs and packed batches on the GPU.

Host-side mirror of the reference's graph construction:

* `generate_edges(struct, inter_radius, intra_radius, prune)` -- drop-in for
  /root/reference/point_vs/preprocessing/preprocessing.py:68-155 (same
  argument meaning, same return triple, same edge ORDER), computed by the K1
  cell-list kernel.
* `radius_graph_batch(...)` -- the packed-batch path: many complexes in one
  launch, destination-sorted CSR + work tiles, ready for the EGNN kernels.
* `csr_from_edge_index(...)` -- any caller-supplied PyG `edge_index`
  (pnn_geometric_base.py:55-58; attribution passes masked, arbitrarily
  ordered edges) -> the same CSR, remembering the caller's order.
* `PackedBatch` -- duck-types the PyG `Batch` the models consume
  (data_loaders.py:381-391, 517-520).
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _cabi
from ._cabi import check, lib, ptr, stream


class CSRGraph:
    """Destination-sorted CSR + work tiles, all on one CUDA device.

    row_ptr [N+1] int32, col [E] int32, attr [E] uint8 (edge class),
    perm [E] int32 or None: perm[p] = index in the caller's edge order of the
    edge stored at CSR slot p.
    """

    def __init__(self, n_nodes, n_edges, row_ptr, col, attr, perm=None,
                 n_inter=None, ref_pos=None, n_classes=3):
        self.n_nodes, self.n_edges = int(n_nodes), int(n_edges)
        self.row_ptr, self.col, self.attr = row_ptr, col, attr
        self.perm, self.n_inter, self.ref_pos = perm, n_inter, ref_pos
        self.n_classes = n_classes
        self.device = row_ptr.device
        self._inv_perm = None
        self._csc = None
        self.exact_edge_count = True
        self.symmetric = False    # set by K1: i-j is an edge iff j-i is
        self.n_edges_dev = None
        self._overflow = None
        # work-tile partitions, built on first use: node-aligned tiles for the
        # FFMA kernels and the backward, edge-packed tiles for tcgen05
        self.tile_ptr = self.n_tiles = None
        self.n_tiles_cap = 0
        self.ptile_last = self.n_ptiles = None
        self.n_ptiles_cap = 0

    def check_overflow(self):
        """Host sync: raise if a capacity-bounded build dropped edges."""
        if self._overflow is not None and int(self._overflow.item()):
            raise _cabi.PvsError(
                f'edge capacity {self.n_edges} too small for '
                f'{int(self.n_edges_dev.item())} edges')

    def true_edge_count(self):
        """Exact number of edges (a host sync for capacity-bounded graphs)."""
        if self.exact_edge_count:
            return self.n_edges
        return int(self.n_edges_dev.item())

    def _exact(self):
        """Trim a capacity-bounded graph to its exact size (host sync)."""
        if not self.exact_edge_count:
            self.check_overflow()
            e = int(self.n_edges_dev.item())
            self.n_edges = e
            self.col, self.attr = self.col[:e], self.attr[:e]
            self.exact_edge_count = True

    def _build_node_tiles(self):
        h = lib()
        dev = self.device
        cap = h.pvs_tiles_capacity(self.n_nodes, self.n_edges)
        self.n_tiles_cap = int(cap)
        self.tile_ptr = torch.empty(cap + 1, dtype=torch.int32, device=dev)
        self.n_tiles = torch.zeros(1, dtype=torch.int32, device=dev)
        scratch = _cabi.scratch(
            'tiles', max(1, int(h.pvs_tiles_scratch_bytes(self.n_nodes))), dev)
        with torch.cuda.device(dev):
            check(h.pvs_build_tiles(ptr(self.row_ptr), self.n_nodes,
                                    ptr(self.tile_ptr), ptr(self.n_tiles),
                                    ptr(scratch), stream()), 'pvs_build_tiles')

    def _build_packed_tiles(self):
        h = lib()
        dev = self.device
        pcap = int(h.pvs_packed_tiles_capacity(self.n_edges))
        self.n_ptiles_cap = pcap
        self.ptile_last = torch.empty(pcap, dtype=torch.int32, device=dev)
        self.n_ptiles = torch.empty(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(h.pvs_build_packed_tiles(
                ptr(self.row_ptr), self.n_nodes, self.n_edges,
                ptr(self.ptile_last), ptr(self.n_ptiles), stream()),
                'pvs_build_packed_tiles')

    def c_struct(self, node_tiles=True, packed_tiles=True):
        """struct pvs_graph.  node_tiles / packed_tiles say which work-tile
        partitions the callee needs (FFMA kernels and every backward: node
        tiles; tcgen05 forward: packed tiles); each is built once."""
        if node_tiles and self.tile_ptr is None:
            self._build_node_tiles()
        if packed_tiles and self.ptile_last is None:
            self._build_packed_tiles()
        return _cabi.Graph(self.n_nodes, self.n_edges, ptr(self.row_ptr),
                           ptr(self.col), ptr(self.attr), ptr(self.tile_ptr),
                           ptr(self.n_tiles), self.n_tiles_cap,
                           ptr(self.ptile_last), ptr(self.n_ptiles),
                           self.n_ptiles_cap)

    # -- orderings -------------------------------------------------------
    def rows(self):
        """Destination node of every CSR edge, int64 [E]."""
        self._exact()
        deg = (self.row_ptr[1:] - self.row_ptr[:-1]).long()
        return torch.repeat_interleave(
            torch.arange(self.n_nodes, device=self.device), deg,
            output_size=self.n_edges)

    def inv_perm(self):
        if self.perm is None:
            return None
        if self._inv_perm is None:
            inv = torch.empty_like(self.perm, dtype=torch.long)
            inv[self.perm.long()] = torch.arange(
                self.n_edges, device=self.device)
            self._inv_perm = inv
        return self._inv_perm

    def to_caller_order(self, per_edge):
        """CSR-ordered per-edge tensor -> the caller's edge order."""
        if self.perm is None:
            return per_edge
        return per_edge.index_select(0, self.inv_perm())

    def from_caller_order(self, per_edge):
        if self.perm is None:
            return per_edge
        return per_edge.index_select(0, self.perm.long())

    def csc(self):
        """(csc_ptr [N+1], csc_eid [E]): CSR edge ids grouped by neighbour."""
        if self._csc is None and self.symmetric:
            # radius graphs are symmetric: csc_ptr IS row_ptr, and the edge ids
            # come from one search pass (no counting sort, 85 -> 10 us).  An
            # edge without a reverse raises the graph's overflow flag.
            csc_eid = torch.empty(max(1, self.n_edges), dtype=torch.int32,
                                  device=self.device)
            if self._overflow is None:
                self._overflow = torch.zeros(1, dtype=torch.int32,
                                             device=self.device)
            g = self.c_struct(node_tiles=False, packed_tiles=False)
            with torch.cuda.device(self.device):
                check(lib().pvs_csr_transpose_symmetric(
                    C.byref(g), ptr(csc_eid), ptr(self._overflow), stream()),
                    'pvs_csr_transpose_symmetric')
            self._csc = (self.row_ptr, csc_eid)
        if self._csc is None:
            # no _exact(): the transpose reads the true edge count on the device
            # (row_ptr[n]), so a capacity-bounded graph stays sync-free
            h = lib()
            dev = self.device
            csc_ptr = torch.empty(self.n_nodes + 1, dtype=torch.int32, device=dev)
            csc_eid = torch.empty(max(1, self.n_edges), dtype=torch.int32,
                                  device=dev)
            nbytes = 2 * ((self.n_nodes * 4 + 255) // 256 * 256) + \
                int(h.pvs_scan_scratch_bytes(self.n_nodes)) + 256
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            g = self.c_struct()
            check(h.pvs_csr_transpose(C.byref(g), ptr(csc_ptr), ptr(csc_eid),
                                      ptr(scratch), stream()),
                  'pvs_csr_transpose')
            self._csc = (csc_ptr, csc_eid)
        return self._csc

    def edge_index(self, order='csr'):
        """PyG-style [2,E] int64 edge_index ('csr' or 'reference' order)."""
        ei = torch.stack([self.rows(), self.col.long()])
        if order == 'reference':
            if self.ref_pos is None:
                raise ValueError('graph was built without ref_pos')
            out = torch.empty_like(ei)
            out[:, self.ref_pos.long()] = ei
            return out
        return ei

    def edge_attr_onehot(self, order='csr'):
        self._exact()
        a = torch.nn.functional.one_hot(self.attr.long(), self.n_classes)
        if order == 'reference':
            out = torch.empty_like(a)
            out[self.ref_pos.long()] = a
            return out
        return a


_INT_CACHE = {}


def _device_ints(arr, device):
    """Small host int32 array -> device tensor, cached by content.  A copy
    from pageable host memory blocks the host until the stream drains, i.e. a
    hidden full sync in every step; batches of a repeated shape (the common
    case in screening) hit the cache instead."""
    key = (str(device), arr.tobytes())
    t = _INT_CACHE.get(key)
    if t is None:
        if len(_INT_CACHE) > 256:
            _INT_CACHE.clear()
        t = torch.from_numpy(arr.copy()).to(device)
        _INT_CACHE[key] = t
    return t


def auto_edge_capacity(n_atoms, inter_radius, intra_radius):
    """Upper bound on the directed edges of `n_atoms` protein-ligand atoms:
    0.375 r^3 neighbours per atom (24 at r = 4 A; heavy-atom + polar-hydrogen
    densities of crystal structures stay below 0.09 atoms / A^3, i.e. below
    0.377 r^3 neighbours in a full sphere, and atoms near the surface of the
    crop have fewer)."""
    r = max(float(inter_radius), float(intra_radius), 1.0)
    return int(math.ceil(0.375 * r ** 3)) * int(n_atoms)


def _to_device(arr, dtype, device):
    """Host array -> device tensor through pinned memory, without blocking:
    a copy from pageable memory would first drain the stream (a hidden sync
    per batch).  Device tensors pass through."""
    if torch.is_tensor(arr) and arr.is_cuda:
        return arr.to(device=device, dtype=dtype)
    t = torch.as_tensor(arr, dtype=dtype)
    if device.type != 'cuda' or t.numel() == 0:
        return t.to(device)
    return t.contiguous().pin_memory().to(device, non_blocking=True)


def radius_graph_batch(coords, bp, complex_ptr, inter_radius=4.0,
                       intra_radius=2.0, with_ref_pos=False, device=None,
                       edge_capacity=None):
    """K1 over a packed batch.

    coords: float64 [N,3] (tensor or array), bp: int [N], complex_ptr: int
    [B+1] host array of node offsets.  Returns a CSRGraph whose within-row
    order is the stable sort by destination of the reference's edge list.

    edge_capacity=None reads the edge count back once (the only host sync) to
    size col/attr exactly.  With an int (an upper bound on E) or 'auto'
    (`auto_edge_capacity`: 24 edges per atom at a 4 A cut-off, growing with
    the cube of the radius) nothing is read back: `n_edges` of the result is
    the capacity, `n_edges_dev` the true count on the device, and
    `check_overflow()` (a sync, call it whenever convenient) raises if the
    bound was too small.  Forward, backward and the CSR transpose all take the
    true count from the device, so a training step on such a graph has no
    host sync at all.
    """
    device = torch.device(device or 'cuda')
    h = lib()
    coords = torch.as_tensor(coords, dtype=torch.float64).to(device).contiguous()
    bp = torch.as_tensor(np.array(bp) if isinstance(bp, np.ndarray) else bp
                         ).to(device=device, dtype=torch.int32).contiguous()
    cptr_host = np.asarray(complex_ptr, dtype=np.int32)
    n = int(coords.shape[0])
    n_complexes = len(cptr_host) - 1
    if n_complexes < 0 or (n_complexes >= 0 and int(cptr_host[-1]) != n):
        raise ValueError('complex_ptr must end at the number of atoms')
    max_n = int(np.max(np.diff(cptr_host))) if n_complexes > 0 else 0
    cptr = _device_ints(cptr_host, device)
    deg = torch.empty(max(1, n), dtype=torch.int32, device=device)
    n_inter = torch.empty(max(1, n), dtype=torch.int32, device=device)
    row_ptr = torch.empty(n + 1, dtype=torch.int32, device=device)
    scratch = _cabi.scratch('k1_scan', int(h.pvs_scan_scratch_bytes(n)) + 256,
                            device)
    # neighbour masks kept between the passes (skipped when they would be huge
    # or when the reference-order positions need the per-complex pass anyway)
    mask_bytes = int(h.pvs_radius_graph_mask_bytes(n, max_n))
    masks = None
    if not with_ref_pos and 0 < mask_bytes <= (1 << 30):
        masks = _cabi.scratch('k1_masks', mask_bytes, device)
    with torch.cuda.device(device):
        check(h.pvs_radius_graph_count(
            ptr(coords), ptr(bp), ptr(cptr), n_complexes, n, max_n,
            C.c_double(inter_radius), C.c_double(intra_radius), ptr(deg),
            ptr(n_inter), ptr(row_ptr), ptr(masks), ptr(scratch), stream()),
            'pvs_radius_graph_count')
        if edge_capacity is None:
            n_edges = int(row_ptr[-1].item())   # the one host sync
            capacity = n_edges
        else:
            capacity = auto_edge_capacity(n, inter_radius, intra_radius) \
                if edge_capacity == 'auto' else int(edge_capacity)
            n_edges = capacity
        col = torch.empty(max(1, capacity), dtype=torch.int32, device=device)
        attr = torch.empty(max(1, capacity), dtype=torch.uint8, device=device)
        ref_pos = torch.empty(max(1, capacity), dtype=torch.int32,
                              device=device) if with_ref_pos else None
        overflow = torch.zeros(1, dtype=torch.int32, device=device)
        check(h.pvs_radius_graph_fill(
            ptr(coords), ptr(bp), ptr(cptr), n_complexes, n, max_n,
            C.c_double(inter_radius), C.c_double(intra_radius), ptr(n_inter),
            ptr(row_ptr), ptr(masks), capacity, ptr(col), ptr(attr),
            ptr(ref_pos), ptr(overflow), stream()),
            'pvs_radius_graph_fill')
        g = CSRGraph(n, n_edges, row_ptr, col[:n_edges], attr[:n_edges],
                     perm=None, n_inter=n_inter[:n],
                     ref_pos=None if ref_pos is None else ref_pos[:n_edges])
    g.complex_ptr = cptr
    g.complex_ptr_host = cptr_host
    g.n_edges_dev = row_ptr[-1:]
    g.exact_edge_count = edge_capacity is None
    g.symmetric = True
    g._overflow = overflow
    return g


def prune_mask(graph):
    """keep[i] for `prune=True` (preprocessing.py:144-153)."""
    h = lib()
    keep = torch.zeros(max(1, graph.n_nodes), dtype=torch.uint8,
                       device=graph.device)
    with torch.cuda.device(graph.device):
        check(h.pvs_prune_mask(ptr(graph.row_ptr), ptr(graph.col),
                               ptr(graph.n_inter), ptr(graph.complex_ptr),
                               len(graph.complex_ptr_host) - 1, graph.n_nodes,
                               ptr(keep), stream()), 'pvs_prune_mask')
    return keep[:graph.n_nodes].bool()


def generate_edges(struct, inter_radius=4.0, intra_radius=2.0, prune=True,
                   synthpharm=False, device=None):
    """Drop-in for the reference's generate_edges (same signature/returns).

    struct: DataFrame with x, y, z, bp columns.  Returns
    (struct, (row, col), edge_attrs) with numpy int64/int32 arrays in the
    reference's order ([inter | intra], each half row-major).  As in the
    reference, `struct` is re-indexed in place and pruned atoms are dropped
    from it.
    """
    struct.reset_index(inplace=True, drop=True)
    if synthpharm:
        struct['bp'] = struct['atom_id'].apply(lambda v: int(v <= 2))
    coords = np.vstack([struct.x.to_numpy(), struct.y.to_numpy(),
                        struct.z.to_numpy()]).T.astype(np.float64)
    bp = struct.bp.to_numpy()
    n = len(coords)
    g = radius_graph_batch(coords, bp, [0, n], inter_radius, intra_radius,
                           with_ref_pos=True, device=device)
    if prune and int(g.n_inter.sum().item()) > 0:
        keep = prune_mask(g).cpu().numpy()
        drop = struct.index[~keep]
        struct.drop(drop, inplace=True)
        return generate_edges(struct.copy(), inter_radius, intra_radius,
                              False, device=device)
    ei = g.edge_index('reference').cpu().numpy()
    attr = torch.empty_like(g.attr)
    attr[g.ref_pos.long()] = g.attr
    return struct, (ei[0], ei[1]), attr.cpu().numpy().astype(np.int32)


def csr_from_edge_index(edge_index, edge_attr, n_nodes):
    """Caller-ordered PyG edge_index [2,E] (+ one-hot edge_attr [E,D] or None)
    -> CSRGraph with `perm`.  Raises IndexError on out-of-range node ids."""
    h = lib()
    device = edge_index.device
    _cabi.require_cuda(edge_index)
    ei = edge_index.to(torch.int64).contiguous()
    n_edges = int(ei.shape[1])
    n_classes = 0
    onehot = None
    if edge_attr is not None:
        if edge_attr.dim() != 2 or edge_attr.shape[0] != n_edges:
            raise ValueError('edge_attr must be [E, n_classes] one-hot')
        n_classes = int(edge_attr.shape[1])
        onehot = edge_attr.to(device=device, dtype=torch.int64).contiguous()
    deg = torch.empty(max(1, n_nodes), dtype=torch.int32, device=device)
    row_ptr = torch.empty(n_nodes + 1, dtype=torch.int32, device=device)
    col = torch.empty(max(1, n_edges), dtype=torch.int32, device=device)
    attr = torch.zeros(max(1, n_edges), dtype=torch.uint8, device=device)
    perm = torch.empty(max(1, n_edges), dtype=torch.int32, device=device)
    bad = torch.zeros(1, dtype=torch.int32, device=device)
    nbytes = (n_nodes * 4 + 255) // 256 * 256 + \
        int(h.pvs_scan_scratch_bytes(n_nodes)) + 256
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        check(h.pvs_edge_index_to_csr(
            ptr(ei), C.c_int64(n_edges), ptr(onehot), n_classes, n_nodes,
            ptr(deg), ptr(row_ptr), ptr(col), ptr(attr), ptr(perm), ptr(bad),
            ptr(scratch), stream()), 'pvs_edge_index_to_csr')
        g = CSRGraph(n_nodes, n_edges, row_ptr, col[:n_edges], attr[:n_edges],
                     perm=perm[:n_edges], n_classes=max(n_classes, 1))
    g._bad = bad
    return g


def dropout_adj(csr, p, force_undirected=True, generator=None):
    """Edge dropout of the reference's training path: `dropout_adj(edges,
    edge_attributes, p, force_undirected=True, training=True)`
    (egnn_satorras.py:320-323; torch_geometric 2.0.4 `utils.dropout_adj`,
    third-party, restated from its published source): keep the edges with
    row < col, drop each with probability p, then add the mirror image of
    every survivor.  -> a new CSRGraph (`perm` = order [kept | mirrored]).
    The random stream is this device's, so results match the reference in
    distribution, not draw for draw."""
    if p < 0.0 or p > 1.0:
        raise ValueError(f'Dropout probability has to be between 0 and 1 (got {p})')
    csr._exact()
    row, col, attr = csr.rows(), csr.col.long(), csr.attr.long()
    if force_undirected:
        keep = row < col
        row, col, attr = row[keep], col[keep], attr[keep]
    prob = torch.full((row.numel(),), 1.0 - p, dtype=torch.float32,
                      device=csr.device)
    mask = torch.bernoulli(prob, generator=generator).to(torch.bool)
    row, col, attr = row[mask], col[mask], attr[mask]
    if force_undirected:
        ei = torch.stack([torch.cat([row, col]), torch.cat([col, row])])
        attr = torch.cat([attr, attr])
    else:
        ei = torch.stack([row, col])
    onehot = torch.nn.functional.one_hot(attr, csr.n_classes)
    out = csr_from_edge_index(ei, onehot, csr.n_nodes)
    out._key_refs = (ei, onehot)
    return out


class PackedBatch:
    """Variable-size complexes packed for the EGNN kernels; duck-types the PyG
    `Batch` consumed by `forward(graph)` (x, pos, edge_index, edge_attr,
    batch, y, lig_fname, rec_fname) and carries the prebuilt CSR as
    `pvs_csr` and the node offsets as `graph_ptr`, so the model neither sorts
    edges nor derives the batch size on the host.  edge_index / edge_attr /
    batch are materialised only if somebody asks for them."""

    def __init__(self, x, pos, csr, graph_ptr, y=None, lig_fname=None,
                 rec_fname=None):
        self.x, self.pos, self.pvs_csr, self.graph_ptr = x, pos, csr, graph_ptr
        self.y = y
        self.num_graphs = int(graph_ptr.numel()) - 1
        self.lig_fname = lig_fname if lig_fname is not None else []
        self.rec_fname = rec_fname if rec_fname is not None else []
        self._edge_index = self._edge_attr = self._batch = None

    @property
    def edge_index(self):
        if self._edge_index is None:
            self._edge_index = self.pvs_csr.edge_index('csr')
        return self._edge_index

    @property
    def edge_attr(self):
        if self._edge_attr is None:
            self._edge_attr = self.pvs_csr.edge_attr_onehot('csr')
        return self._edge_attr

    @property
    def batch(self):
        if self._batch is None:
            sizes = (self.graph_ptr[1:] - self.graph_ptr[:-1]).long()
            self._batch = torch.repeat_interleave(
                torch.arange(self.num_graphs, device=sizes.device), sizes,
                output_size=int(self.x.shape[0]))
        return self._batch

    def to(self, device):
        return self

    @staticmethod
    def from_arrays(coords, bp, feats, complex_ptr, inter_radius=4.0,
                    intra_radius=4.0, y=None, device=None, lig_fname=None,
                    rec_fname=None, edge_capacity=None):
        """coords f64 [N,3], bp [N], feats f32 [N,F], complex_ptr [B+1]
        (host arrays, or device tensors for coords/bp/feats)."""
        device = torch.device(device or 'cuda')
        coords_d = _to_device(coords, torch.float64, device)
        if isinstance(bp, np.ndarray):
            bp = _to_device(bp, torch.int32, device)
        csr = radius_graph_batch(coords_d, bp, complex_ptr, inter_radius,
                                 intra_radius, device=device,
                                 edge_capacity=edge_capacity)
        x = _to_device(feats, torch.float32, device)
        return PackedBatch(x, coords_d.float(), csr, csr.complex_ptr, y=y,
                           lig_fname=lig_fname, rec_fname=rec_fname)
