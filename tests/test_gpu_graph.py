"""K1 (radius graph), CSR conversion and tile partition on the GPU, through
the C ABI, against the reference goldens and the CPU oracle.  Bit-exact."""
import numpy as np
import pandas as pd
import pytest
import torch

from tests import helpers
from tests.test_oracle_golden import (TOY_COORDS, TOY_BP, TOY_ROW, TOY_COL,
                                      TOY_ATTR, TOY_PRUNED_ROW,
                                      TOY_PRUNED_COL, TOY_PRUNED_ATTR,
                                      EDGE_CASES)

pytestmark = pytest.mark.gpu


def _struct(coords, bp):
    return pd.DataFrame({'x': coords[:, 0], 'y': coords[:, 1],
                         'z': coords[:, 2], 'bp': bp})


def test_toy_golden_reference_order():
    from pointvs_b200 import generate_edges
    _, (row, col), attr = generate_edges(
        _struct(TOY_COORDS, TOY_BP), 2.1, 1.1, prune=False)
    np.testing.assert_array_equal(row, TOY_ROW)
    np.testing.assert_array_equal(col, TOY_COL)
    np.testing.assert_array_equal(attr, TOY_ATTR)


def test_toy_golden_prune():
    from pointvs_b200 import generate_edges
    out, (row, col), attr = generate_edges(
        _struct(TOY_COORDS, TOY_BP), 2.1, 1.1, prune=True)
    np.testing.assert_array_equal(row, TOY_PRUNED_ROW)
    np.testing.assert_array_equal(col, TOY_PRUNED_COL)
    np.testing.assert_array_equal(attr, TOY_PRUNED_ATTR)
    assert len(out) == 8


@pytest.mark.parametrize('name', EDGE_CASES)
def test_generate_edges_vs_reference_golden(name):
    from pointvs_b200 import generate_edges
    g = helpers.load_npz('edges.npz')
    inter, intra = g[name + '.radii']
    _, (row, col), attr = generate_edges(
        _struct(g[name + '.coords'], g[name + '.bp']), inter, intra,
        prune=False)
    np.testing.assert_array_equal(row, g[name + '.row'])
    np.testing.assert_array_equal(col, g[name + '.col'])
    np.testing.assert_array_equal(attr, g[name + '.attr'])


def test_prune_vs_reference_golden():
    from pointvs_b200 import generate_edges
    g = helpers.load_npz('edges.npz')
    inter, intra = g['prune.radii']
    out, (row, col), attr = generate_edges(
        _struct(g['prune.coords'], g['prune.bp']), inter, intra, prune=True)
    assert len(out) == int(g['prune.n_kept'][0])
    np.testing.assert_array_equal(row, g['prune.row'])
    np.testing.assert_array_equal(col, g['prune.col'])
    np.testing.assert_array_equal(attr, g['prune.attr'])


@pytest.mark.parametrize('radii', [(4.0, 4.0), (4.0, 2.0), (7.5, 3.0)])
def test_packed_batch_vs_oracle(radii):
    """Ragged batch: per-complex edge lists must equal the oracle's, and the
    CSR must be the stable sort by destination of the reference order."""
    from oracle import radius_graph as rg
    from pointvs_b200.graph import radius_graph_batch
    from pointvs_b200.synthetic import synthetic_batch
    coords, bp, _, cptr = synthetic_batch(100, 6, n_atoms=500, n_lig=25,
                                          ragged=True)
    g = radius_graph_batch(coords, bp, cptr, *radii, with_ref_pos=True)
    rows, cols, attrs, csr_rows, csr_cols, csr_attrs = [], [], [], [], [], []
    for c in range(len(cptr) - 1):
        lo, hi = cptr[c], cptr[c + 1]
        r, cc, a = rg.radius_graph_c(coords[lo:hi], bp[lo:hi], *radii)
        order = rg.csr_order(r, cc, a)
        rows.append(r + lo); cols.append(cc + lo); attrs.append(a)
        csr_rows.append(r[order] + lo); csr_cols.append(cc[order] + lo)
        csr_attrs.append(a[order])
    want_row, want_col, want_attr = map(np.concatenate, (rows, cols, attrs))
    assert g.n_edges == len(want_row)
    # CSR order
    ei = g.edge_index('csr').cpu().numpy()
    np.testing.assert_array_equal(ei[0], np.concatenate(csr_rows))
    np.testing.assert_array_equal(ei[1], np.concatenate(csr_cols))
    np.testing.assert_array_equal(g.attr.cpu().numpy(),
                                  np.concatenate(csr_attrs))
    # reference (PyG-collated) order through ref_pos
    ei = g.edge_index('reference').cpu().numpy()
    np.testing.assert_array_equal(ei[0], want_row)
    np.testing.assert_array_equal(ei[1], want_col)
    ref_attr = torch.empty_like(g.attr)
    ref_attr[g.ref_pos.long()] = g.attr
    np.testing.assert_array_equal(ref_attr.cpu().numpy(), want_attr)


def test_config3_shape_properties():
    """Full-size complexes (16 x 1000 atoms): symmetry of the edge multiset,
    degree statistics, and agreement with the oracle on one complex."""
    from oracle import radius_graph as rg
    from pointvs_b200.graph import radius_graph_batch
    from pointvs_b200.synthetic import synthetic_batch
    coords, bp, _, cptr = synthetic_batch(0, 16)
    g = radius_graph_batch(coords, bp, cptr, 4.0, 4.0)
    ei = g.edge_index('csr').cpu().numpy()
    attr = g.attr.cpu().numpy()
    fwd = np.lexsort((attr, ei[1], ei[0]))
    rev = np.lexsort((attr, ei[0], ei[1]))
    np.testing.assert_array_equal(ei[0][fwd], ei[1][rev])
    np.testing.assert_array_equal(ei[1][fwd], ei[0][rev])
    np.testing.assert_array_equal(attr[fwd], attr[rev])
    assert 14.5 < g.n_edges / 16000 < 16.0
    r, c, a = rg.radius_graph_c(coords[:1000], bp[:1000], 4.0, 4.0)
    want = rg.canonical(r, c, a)
    sel = ei[0] < 1000
    got = rg.canonical(ei[0][sel], ei[1][sel], attr[sel])
    for w, x in zip(want, got):
        np.testing.assert_array_equal(w, x)


def test_degenerate_inputs():
    from pointvs_b200.graph import radius_graph_batch
    g = radius_graph_batch(np.zeros((0, 3)), np.zeros(0), [0], 4.0, 2.0)
    assert g.n_edges == 0 and g.n_nodes == 0
    g = radius_graph_batch(np.zeros((1, 3)), np.ones(1), [0, 1], 4.0, 2.0)
    assert g.n_edges == 0
    # empty complex between two real ones, coincident atoms
    coords = np.array([[0, 0, 0], [0, 0, 0], [1, 0, 0], [50, 0, 0], [51, 0, 0]],
                      dtype=np.float64)
    bp = np.array([0, 1, 1, 0, 1])
    g = radius_graph_batch(coords, bp, [0, 3, 3, 5], 4.0, 2.0)
    ei = g.edge_index('csr').cpu().numpy()
    want = {(0, 2, 1), (0, 2, 0), (1, 2, 2), (2, 0, 1), (2, 0, 0), (2, 1, 2),
            (3, 4, 1), (3, 4, 0), (4, 3, 1), (4, 3, 0)}
    got = set(zip(ei[0].tolist(), ei[1].tolist(), g.attr.cpu().tolist()))
    assert got == want and g.n_edges == len(want)


def test_large_complex_many_cells():
    """A 6000-atom complex spanning far more than 16 cells per axis."""
    from oracle import radius_graph as rg
    from pointvs_b200.graph import radius_graph_batch
    rng = np.random.default_rng(5)
    coords = rng.uniform(-60, 60, size=(6000, 3)).astype(np.float32).astype(
        np.float64)
    coords[:3000] = rng.uniform(-8, 8, size=(3000, 3)).astype(np.float32)
    bp = (rng.random(6000) < 0.9).astype(np.int32)
    g = radius_graph_batch(coords, bp, [0, 6000], 4.0, 2.0)
    r, c, a = rg.radius_graph_c(coords, bp, 4.0, 2.0)
    order = rg.csr_order(r, c, a)
    ei = g.edge_index('csr').cpu().numpy()
    np.testing.assert_array_equal(ei[0], r[order])
    np.testing.assert_array_equal(ei[1], c[order])
    np.testing.assert_array_equal(g.attr.cpu().numpy(), a[order])


def test_csr_from_shuffled_edge_index_is_stable():
    from pointvs_b200.graph import csr_from_edge_index
    g = helpers.load_npz('fixture82.npz')
    ei = torch.from_numpy(g['edge_index']).long()
    ea = torch.from_numpy(g['edge_attr']).long()
    perm = torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(0))
    ei_s, ea_s = ei[:, perm].cuda(), ea[perm].cuda()
    csr = csr_from_edge_index(ei_s, ea_s, 164)
    order = np.argsort(ei_s[0].cpu().numpy(), kind='stable')
    np.testing.assert_array_equal(csr.perm.cpu().numpy(), order)
    np.testing.assert_array_equal(csr.col.cpu().numpy(),
                                  ei_s[1].cpu().numpy()[order])
    np.testing.assert_array_equal(csr.attr.cpu().numpy(),
                                  ea_s.argmax(1).cpu().numpy()[order])
    deg = np.bincount(ei_s[0].cpu().numpy(), minlength=164)
    np.testing.assert_array_equal(np.diff(csr.row_ptr.cpu().numpy()), deg)
    # round trip of a per-edge tensor
    vals = torch.arange(ei.shape[1], device='cuda', dtype=torch.float32)
    assert torch.equal(csr.to_caller_order(csr.from_caller_order(vals)), vals)
    # CSC transpose groups the same edges by neighbour
    csc_ptr, csc_eid = csr.csc()
    cols = csr.col.cpu().numpy()
    eid = csc_eid.cpu().numpy()[:csr.n_edges]
    np.testing.assert_array_equal(eid, np.argsort(cols, kind='stable'))
    np.testing.assert_array_equal(np.diff(csc_ptr.cpu().numpy()),
                                  np.bincount(cols, minlength=164))


def test_out_of_range_edge_index_raises():
    from pointvs_b200.egnn import _csr_for
    ei = torch.tensor([[0, 1, 7], [1, 0, 2]], device='cuda')
    with pytest.raises(IndexError):
        _csr_for(ei, None, 3)


@pytest.mark.parametrize('radii', [(4.0, 4.0), (4.0, 2.0)])
def test_tile_partition(radii):
    from pointvs_b200.graph import radius_graph_batch
    from pointvs_b200.synthetic import synthetic_batch
    coords, bp, _, cptr = synthetic_batch(7, 5, ragged=True)
    g = radius_graph_batch(coords, bp, cptr, *radii)
    g.c_struct()                      # tile partitions are built on first use
    t = int(g.n_tiles.item())
    assert 0 < t <= g.n_tiles_cap
    tp = g.tile_ptr.cpu().numpy()[:t + 1]
    rp = g.row_ptr.cpu().numpy()
    assert tp[0] == 0 and tp[-1] == g.n_nodes and np.all(np.diff(tp) > 0)
    edges = rp[tp[1:]] - rp[tp[:-1]]
    nodes = np.diff(tp)
    assert np.all((edges <= 128) | (nodes == 1)) and np.all(nodes <= 128)
    # greedy: neighbouring tiles cannot be merged (inside one 256-node chunk)
    same_chunk = (tp[1:-1] % 256) != 0
    merged_e = edges[:-1] + edges[1:]
    merged_n = nodes[:-1] + nodes[1:]
    assert np.all(~same_chunk | (merged_e > 128) | (merged_n > 128))
    # edge-packed tiles of the tcgen05 kernel: tile t owns edges
    # [128 t, 128 t + 128); last[t] is the node of its last edge
    e = int(rp[-1])
    tpk = int(g.n_ptiles.item())
    assert tpk == max(1, -(-e // 128)) and tpk <= g.n_ptiles_cap
    last = g.ptile_last.cpu().numpy()[:tpk]
    row_of = np.repeat(np.arange(g.n_nodes), np.diff(rp))
    want = row_of[np.minimum(128 * (np.arange(tpk) + 1), e) - 1]
    np.testing.assert_array_equal(last, want)


def test_high_degree_node_gets_own_tile():
    from pointvs_b200.graph import csr_from_edge_index
    n = 400
    hub = torch.zeros(n - 1, dtype=torch.long)
    others = torch.arange(1, n)
    ei = torch.cat([torch.stack([hub, others]), torch.stack([others, hub])], 1)
    csr = csr_from_edge_index(ei.cuda(), None, n)
    csr.c_struct()
    t = int(csr.n_tiles.item())
    tp = csr.tile_ptr.cpu().numpy()[:t + 1]
    assert tp[0] == 0 and tp[1] == 1 and tp[-1] == n


def test_batch_to_ptr():
    from pointvs_b200.dense import batch_to_ptr
    batch = torch.tensor([0, 0, 0, 2, 2, 5], device='cuda')
    ptr = batch_to_ptr(batch, 7).cpu().tolist()
    assert ptr == [0, 3, 3, 5, 5, 5, 6, 6]
    assert batch_to_ptr(torch.zeros(0, dtype=torch.long, device='cuda'),
                        2).cpu().tolist() == [0, 0, 0]


def test_capacity_bounded_build_matches_exact_build():
    """No-host-sync path: col/attr sized by an upper bound."""
    from pointvs_b200._cabi import PvsError
    from pointvs_b200.graph import radius_graph_batch
    from pointvs_b200.synthetic import synthetic_batch
    coords, bp, _, cptr = synthetic_batch(40, 4, 600, 20, ragged=True)
    exact = radius_graph_batch(coords, bp, cptr, 4.0, 4.0)
    loose = radius_graph_batch(coords, bp, cptr, 4.0, 4.0, edge_capacity='auto')
    assert loose.n_edges == 24 * exact.n_nodes and not loose.exact_edge_count
    loose.check_overflow()
    assert loose.true_edge_count() == exact.n_edges
    e = exact.n_edges
    assert torch.equal(loose.row_ptr, exact.row_ptr)
    assert torch.equal(loose.col[:e], exact.col)
    assert torch.equal(loose.attr[:e], exact.attr)
    assert torch.equal(loose.edge_index('csr'), exact.edge_index('csr'))
    assert loose.n_edges == e           # trimmed by the call above
    tight = radius_graph_batch(coords, bp, cptr, 4.0, 4.0, edge_capacity=e - 5)
    with pytest.raises(PvsError):
        tight.check_overflow()


def test_mask_reuse_and_recompute_fill_agree():
    """fill from stored masks (packed path) == per-complex fill (ref_pos path)."""
    from pointvs_b200.graph import radius_graph_batch
    from pointvs_b200.synthetic import synthetic_batch
    coords, bp, _, cptr = synthetic_batch(60, 5, 400, 20, ragged=True)
    a = radius_graph_batch(coords, bp, cptr, 4.0, 2.0)
    b = radius_graph_batch(coords, bp, cptr, 4.0, 2.0, with_ref_pos=True)
    assert a.n_edges == b.n_edges
    assert torch.equal(a.row_ptr, b.row_ptr)
    assert torch.equal(a.col, b.col) and torch.equal(a.attr, b.attr)


def test_symmetric_transpose_matches_counting_sort():
    """K1 graphs are symmetric: pvs_csr_transpose_symmetric (csc_ptr = row_ptr,
    reverse edge found by a scan) groups exactly the edges the general
    counting-sort transpose groups, for exact and capacity-bounded graphs."""
    import torch
    from tests import gpu_helpers as gh
    for cap in (None, 'auto'):
        csr = gh.synthetic_graph(321, 5, 700, 25, ragged=True,
                                 edge_capacity=cap).pvs_csr
        assert csr.symmetric
        ptr_s, eid_s = csr.csc()
        csr._csc, csr.symmetric = None, False
        ptr_g, eid_g = csr.csc()
        e = int(csr.true_edge_count())
        assert torch.equal(ptr_s.cpu(), ptr_g.cpu())
        assert int(ptr_g[-1]) == e
        grp = torch.repeat_interleave(
            torch.arange(csr.n_nodes, device='cuda'),
            (ptr_g[1:] - ptr_g[:-1]).long())
        key = grp * (e + 1) + eid_s[:e].long()
        assert torch.equal(eid_s[:e][torch.argsort(key)].cpu(), eid_g[:e].cpu())
        # every edge id appears once, and it is the reverse edge
        assert torch.equal(torch.sort(eid_s[:e]).values.cpu(),
                           torch.arange(e, dtype=torch.int32))
        col = csr.col[:e].long()
        rows = torch.repeat_interleave(
            torch.arange(csr.n_nodes, device='cuda'),
            (csr.row_ptr[1:] - csr.row_ptr[:-1]).long())
        assert torch.equal(col[eid_s[:e].long()], rows)
        assert int(csr._overflow.item()) == 0
