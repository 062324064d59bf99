"""TEST INFRASTRUCTURE ONLY -- CPU restatement of PointVS `generate_edges`.

Follows /root/reference/point_vs/preprocessing/preprocessing.py:68-155 on bare
numpy arrays (coords float64 [N,3], bp int [N]) instead of a DataFrame:

  :108      distances = cdist(coords, coords)  -> _pairwise (un-fused fp64
            sqrt((dx*dx + dy*dy) + dz*dz); scipy 1.7.3 `cdist`, third-party,
            restated; pinned bitwise against scipy in tests/test_oracle_golden)
  :110-117  inter list: d < inter_radius and d > 1e-7, bp differs, row-major
  :119-121  intra list: d < intra_radius and d > 1e-7, NOT filtered by molecule
  :129-135  attrs: inter -> 1 for a (0,1)/(1,0) pair; intra -> 2 if both bp==1
  :137-142  output = concat(inter, intra)
  :144-153  prune: component of edge_indices[0][0], drop the rest, renumber,
            rebuild once without prune

`radius_graph_c` is the same algorithm compiled from oracle/radius_graph.c
(built by oracle/Makefile into oracle/_build/); used for larger parity cases
and as the graph-builder leg of bench.py's cpu_baseline.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _pairwise(coords):
    c = np.ascontiguousarray(coords, dtype=np.float64)
    dx = c[:, None, 0] - c[None, :, 0]
    dy = c[:, None, 1] - c[None, :, 1]
    dz = c[:, None, 2] - c[None, :, 2]
    return np.sqrt((dx * dx + dy * dy) + dz * dz)


def _edge_lists(coords, bp, inter_radius, intra_radius):
    d = _pairwise(coords)
    bp = np.asarray(bp).astype(np.int64)
    ri, ci = np.where((d < inter_radius) & (d > 1e-7))
    keep = bp[ri] != bp[ci]
    ri, ci = ri[keep], ci[keep]
    ra, ca = np.where((d < intra_radius) & (d > 1e-7))
    a_inter = np.zeros(len(ri), dtype=np.int32)
    a_inter[((bp[ri] == 0) & (bp[ci] == 1)) | ((bp[ri] == 1) & (bp[ci] == 0))] = 1
    a_intra = np.zeros(len(ra), dtype=np.int32)
    a_intra[(bp[ra] == 1) & (bp[ca] == 1)] = 2
    row = np.concatenate([ri, ra]).astype(np.int64)
    col = np.concatenate([ci, ca]).astype(np.int64)
    return row, col, np.concatenate([a_inter, a_intra]), len(ri)


def radius_graph(coords, bp, inter_radius=4.0, intra_radius=2.0, prune=False):
    """Returns (kept_node_indices, row, col, attr) in the reference's order
    ([inter | intra], each half row-major)."""
    coords = np.asarray(coords, dtype=np.float64)
    bp = np.asarray(bp)
    kept = np.arange(len(coords))
    row, col, attr, n_inter = _edge_lists(coords, bp, inter_radius, intra_radius)
    if prune and n_inter:
        nbrs = [[] for _ in range(len(coords))]
        for r, c in zip(row.tolist(), col.tolist()):
            nbrs[r].append(c)
        seen = np.zeros(len(coords), dtype=bool)
        stack = [int(row[0])]
        seen[stack[0]] = True
        while stack:
            s = stack.pop()
            for t in nbrs[s]:
                if not seen[t]:
                    seen[t] = True
                    stack.append(t)
        kept = np.nonzero(seen)[0]
        row, col, attr, _ = _edge_lists(
            coords[kept], bp[kept], inter_radius, intra_radius)
    return kept, row, col, attr


def canonical(row, col, attr):
    """Sort the multiset {(row, col, attr)} for order-free comparison."""
    row, col, attr = (np.asarray(a).astype(np.int64) for a in (row, col, attr))
    order = np.lexsort((attr, col, row))
    return row[order], col[order], attr[order]


def csr_order(row, col, attr):
    """Stable sort by destination: the within-row order the CUDA builder emits
    (inter edges by col, then intra edges by col)."""
    order = np.argsort(np.asarray(row), kind='stable')
    return order


_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, '_build', 'libpvs_oracle.so')
        if not os.path.exists(path):
            raise RuntimeError(
                'oracle C library not built; run `make -C oracle` '
                '(or __graft_entry__.build())')
        lib = ctypes.CDLL(path)
        lib.pvs_oracle_radius_graph.restype = ctypes.c_long
        lib.pvs_oracle_radius_graph.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_double,
            ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_long]
        _LIB = lib
    return _LIB


def radius_graph_c(coords, bp, inter_radius=4.0, intra_radius=2.0):
    """C build of the same algorithm (no prune).  Returns (row, col, attr)."""
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    bp32 = np.ascontiguousarray(bp, dtype=np.int32)
    n = len(coords)
    lib = _lib()
    need = lib.pvs_oracle_radius_graph(
        coords.ctypes.data, bp32.ctypes.data, n, inter_radius, intra_radius,
        None, None, None, 0)
    row = np.empty(need, dtype=np.int64)
    col = np.empty(need, dtype=np.int64)
    attr = np.empty(need, dtype=np.int32)
    got = lib.pvs_oracle_radius_graph(
        coords.ctypes.data, bp32.ctypes.data, n, inter_radius, intra_radius,
        row.ctypes.data, col.ctypes.data, attr.ctypes.data, need)
    assert got == need
    return row, col, attr
