"""BASELINE configs[2] at full size (128 x 1000-atom complexes, 8 layers x 64
channels), where the CPU oracle is too slow to run: size-independent
properties of the radius graph and of the scoring pass, per arithmetic mode.

  * graph: CSR well-formed, destination-sorted, symmetric as a set of
    (i, j, class) triples, no self loops, no edge across complexes, squared
    distances inside / outside the radius as fp64 recomputes them;
  * E(3): rotating / translating / reflecting every complex (graph held fixed)
    leaves the scores unchanged;
  * batch independence: permuting the complexes of the batch permutes the
    scores bit for bit (each complex only ever sees its own atoms);
  * a sample of eight complexes scored alone by the CPU oracle.
"""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import bench
from tests import helpers
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu

N_COMPLEXES, N_ATOMS = 128, 1000


@pytest.fixture(scope='module')
def workload():
    import pointvs_b200 as pv
    from pointvs_b200.synthetic import synthetic_batch
    coords, bp, feats, cptr = synthetic_batch(0, N_COMPLEXES, N_ATOMS, 30)
    batch = pv.PackedBatch.from_arrays(coords, bp, feats, cptr,
                                       bench.EDGE_RADIUS, bench.EDGE_RADIUS)
    model = gh.build_model(dict(bench.MODEL_KW), seed=0, coord_gain=1.0)
    model.set_record_side_channels(False)
    model.record_embed_coords = False
    return SimpleNamespace(coords=coords, bp=bp, feats=feats, cptr=cptr,
                           batch=batch, model=model, pos0=batch.pos.clone())


def test_full_size_graph_properties(workload):
    w = workload
    csr = w.batch.pvs_csr
    rp = csr.row_ptr.cpu().numpy().astype(np.int64)
    col = csr.col.cpu().numpy().astype(np.int64)
    attr = csr.attr.cpu().numpy()
    n = N_COMPLEXES * N_ATOMS
    assert rp[0] == 0 and rp[-1] == len(col) == csr.n_edges and len(rp) == n + 1
    assert np.all(np.diff(rp) >= 0)
    row = np.repeat(np.arange(n), np.diff(rp))
    assert np.all(row != col)                                   # no self loops
    assert np.all(row // N_ATOMS == col // N_ATOMS)             # per complex
    # symmetric as a multiset of (i, j, class)
    key = (row * n + col) * 4 + attr
    mirror = (col * n + row) * 4 + attr
    np.testing.assert_array_equal(np.sort(key), np.sort(mirror))
    # classes follow the molecule flags (preprocessing.py:129-135)
    bi, bj = w.bp[row], w.bp[col]
    assert np.all((attr == 2) <= ((bi == 1) & (bj == 1)))
    assert np.all((attr == 1) == ((bi != bj) & (attr != 0)))
    # every listed pair is inside the radius, in fp64
    d = np.linalg.norm(w.coords[row] - w.coords[col], axis=1)
    assert d.max() < bench.EDGE_RADIUS and d.min() > 1e-7
    # and the count matches an independent fp64 count on four complexes
    for c in (0, 37, 90, 127):
        xyz = w.coords[c * N_ATOMS:(c + 1) * N_ATOMS]
        dd = np.sqrt(((xyz[:, None] - xyz[None]) ** 2).sum(-1))
        close = (dd < bench.EDGE_RADIUS) & (dd > 1e-7)
        b = w.bp[c * N_ATOMS:(c + 1) * N_ATOMS]
        cross = b[:, None] != b[None]
        want = int(close.sum() + (close & cross).sum())   # intra list + inter list
        got = int(rp[(c + 1) * N_ATOMS] - rp[c * N_ATOMS])
        assert got == want


def _scores(w, math, pos=None):
    w.model.set_math(math)
    w.batch.pos = (w.pos0 if pos is None else pos).clone()
    with torch.no_grad():
        return w.model(w.batch).reshape(-1).clone()


@pytest.mark.parametrize('math,tol', [('bf16x3', 2e-5), ('fp16x2', 2e-5),
                                      ('fp32', 2e-5)])
def test_full_size_e3_invariance(workload, math, tol):
    w = workload
    base = _scores(w, math)
    assert torch.isfinite(base).all() and base.std() > 1e-4
    q, _ = np.linalg.qr(np.random.default_rng(5).normal(size=(3, 3)))
    rot = torch.from_numpy(q.astype(np.float32)).cuda()      # may include a reflection
    shift = torch.tensor([7.0, -3.0, 11.0]).cuda()
    moved = _scores(w, math, w.pos0 @ rot + shift)
    scale = float(base.abs().max())
    assert float((moved - base).abs().max()) <= tol * scale + 1e-6


@pytest.mark.parametrize('math', ['bf16x3', 'fp16x2'])
def test_full_size_batch_independence_and_determinism(workload, math):
    import pointvs_b200 as pv
    w = workload
    base = _scores(w, math)
    again = _scores(w, math)
    assert torch.equal(base, again)
    perm = np.random.default_rng(9).permutation(N_COMPLEXES)
    idx = (perm[:, None] * N_ATOMS + np.arange(N_ATOMS)[None]).reshape(-1)
    shuffled = pv.PackedBatch.from_arrays(
        w.coords[idx], w.bp[idx], w.feats[idx], w.cptr, bench.EDGE_RADIUS,
        bench.EDGE_RADIUS)
    w.model.set_math(math)
    with torch.no_grad():
        got = w.model(shuffled).reshape(-1)
    # tile boundaries move with the permutation, so the split-node partial
    # sums add in another association: equal to fp32 round-off, not bitwise
    want = base[torch.from_numpy(perm).cuda()]
    assert float((got - want).abs().max()) <= 2e-6 * float(want.abs().max()) + 1e-7


@pytest.mark.parametrize('math,tol', [('bf16x3', 1e-4), ('fp16x2', 1e-4),
                                      ('fp32', 1e-4)])
def test_full_size_sample_against_oracle(workload, math, tol):
    """Eight of the 128 complexes, scored alone by the CPU oracle on the edge
    list the device built, against their rows of the full-batch scores."""
    w = workload
    base = _scores(w, math).cpu().numpy()
    csr = w.batch.pvs_csr
    ei = csr.edge_index('csr').cpu()
    ea = csr.edge_attr_onehot('csr').cpu()
    kw = dict(bench.MODEL_KW)
    for c in (0, 19, 38, 57, 76, 95, 114, 127):
        lo, hi = c * N_ATOMS, (c + 1) * N_ATOMS
        sel = (ei[0] >= lo) & (ei[0] < hi)
        g = SimpleNamespace(
            x=torch.from_numpy(w.feats[lo:hi]), pos=w.pos0[lo:hi].cpu(),
            edge_index=ei[:, sel] - lo, edge_attr=ea[sel],
            batch=torch.zeros(N_ATOMS, dtype=torch.long))
        want, _ = gh.oracle_forward(w.model, kw, g)
        want = float(want.reshape(-1)[0])
        assert abs(base[c] - want) <= tol * abs(want) + 1e-7, (c, base[c], want)
