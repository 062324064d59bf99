"""Pins the oracle to the reference: every comparison here is against
fixtures produced by the UNMODIFIED reference (tests/golden/make_golden.py)
or against golden vectors copied from the reference's own tests."""
import numpy as np
import pytest
import torch

from oracle import radius_graph as rg
from tests import helpers

# /root/reference/test/test_preprocessing_fns.py:16-23 (12-atom toy structure)
TOY_COORDS = np.array(
    [[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 2], [1, 0, 2],
     [0, 1, 2], [1, 1, 2], [0, 0, 6], [1, 0, 6], [0, 1, 6], [1, 1, 6]],
    dtype=np.float64)
TOY_BP = np.array([0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1])
# test_preprocessing_fns.py:32-50
TOY_ROW = [0, 1, 2, 3, 4, 5, 6, 7, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6,
           7, 7, 8, 8, 9, 9, 10, 10, 11, 11]
TOY_COL = [4, 5, 6, 7, 0, 1, 2, 3, 1, 2, 0, 3, 0, 3, 1, 2, 5, 6, 4, 7, 4, 7,
           5, 6, 9, 10, 8, 11, 8, 11, 9, 10]
TOY_ATTR = [1] * 8 + [0] * 8 + [2] * 16
# test_preprocessing_fns.py:53-71
TOY_PRUNED_ROW = TOY_ROW[:24]
TOY_PRUNED_COL = TOY_COL[:24]
TOY_PRUNED_ATTR = TOY_ATTR[:24]


def test_toy_generate_edges_golden():
    kept, row, col, attr = rg.radius_graph(TOY_COORDS, TOY_BP, 2.1, 1.1)
    np.testing.assert_array_equal(row, TOY_ROW)
    np.testing.assert_array_equal(col, TOY_COL)
    np.testing.assert_array_equal(attr, TOY_ATTR)
    assert len(kept) == 12


def test_toy_generate_edges_prune_golden():
    kept, row, col, attr = rg.radius_graph(
        TOY_COORDS, TOY_BP, 2.1, 1.1, prune=True)
    np.testing.assert_array_equal(row, TOY_PRUNED_ROW)
    np.testing.assert_array_equal(col, TOY_PRUNED_COL)
    np.testing.assert_array_equal(attr, TOY_PRUNED_ATTR)
    np.testing.assert_array_equal(kept, np.arange(8))


def test_toy_duplicate_inter_edges():
    # equal radii -> ligand-receptor pairs appear twice
    _, row, col, attr = rg.radius_graph(TOY_COORDS, TOY_BP, 2.1, 2.1)
    assert len(row) == 52
    node0 = sorted(zip(col[row == 0].tolist(), attr[row == 0].tolist()))
    assert node0 == [(1, 0), (2, 0), (3, 0), (4, 0), (4, 1)]


EDGE_CASES = ['syn300_r4_r4', 'syn300_r4_r2', 'syn257_r6_r2', 'syn64_r10_r3',
              'edgecases']


@pytest.mark.parametrize('name', EDGE_CASES)
@pytest.mark.parametrize('impl', ['numpy', 'c'])
def test_radius_graph_vs_reference_golden(name, impl):
    g = helpers.load_npz('edges.npz')
    inter, intra = g[name + '.radii']
    if impl == 'numpy':
        _, row, col, attr = rg.radius_graph(
            g[name + '.coords'], g[name + '.bp'], inter, intra)
    else:
        row, col, attr = rg.radius_graph_c(
            g[name + '.coords'], g[name + '.bp'], inter, intra)
    # exact, including the reference's output ORDER
    np.testing.assert_array_equal(row, g[name + '.row'])
    np.testing.assert_array_equal(col, g[name + '.col'])
    np.testing.assert_array_equal(attr, g[name + '.attr'])


def test_radius_graph_prune_vs_reference_golden():
    g = helpers.load_npz('edges.npz')
    inter, intra = g['prune.radii']
    kept, row, col, attr = rg.radius_graph(
        g['prune.coords'], g['prune.bp'], inter, intra, prune=True)
    assert len(kept) == int(g['prune.n_kept'][0])
    np.testing.assert_array_equal(row, g['prune.row'])
    np.testing.assert_array_equal(col, g['prune.col'])
    np.testing.assert_array_equal(attr, g['prune.attr'])


def test_pairwise_matches_scipy_cdist_bitwise():
    from scipy.spatial.distance import cdist
    from pointvs_b200.synthetic import synthetic_complex
    coords, _, _ = synthetic_complex(3, 400, 20)
    coords = coords + np.random.default_rng(0).normal(size=coords.shape) * 1e-3
    assert np.array_equal(rg._pairwise(coords), cdist(coords, coords))


def test_empty_and_single_atom():
    _, row, col, attr = rg.radius_graph(np.zeros((0, 3)), np.zeros(0), 4, 2)
    assert len(row) == len(col) == len(attr) == 0
    _, row, col, attr = rg.radius_graph(np.zeros((1, 3)), np.ones(1), 4, 2)
    assert len(row) == 0


@pytest.mark.parametrize('name', sorted(helpers.MODEL_GOLDENS))
def test_model_oracle_vs_reference_golden(name):
    """fp32 oracle vs fp32 reference run: same op sequence, so agreement is
    at rounding level (1e-5 relative is a generous bound)."""
    _, _, tasks = helpers.MODEL_GOLDENS[name]
    for task in tasks:
        trace = []
        g, out, x = helpers.run_oracle(name, task=task, trace=trace)
        ref = g[f'out.{task}']
        assert out.shape == tuple(ref.shape) or out.numel() == ref.size
        assert helpers.rel_err(out.detach().numpy().reshape(-1),
                               ref.reshape(-1)) < 1e-5
        n_layers = len(trace)
        for i, rec in enumerate(trace):
            li = i + 1      # golden layer 0 is the embedding
            assert helpers.scaled_err(rec['h'].numpy(),
                                      g[f'layer{li}.h']) < 1e-5
            assert helpers.scaled_err(rec['x'].numpy(),
                                      g[f'layer{li}.x']) < 1e-6
            if f'layer{li}.att' in g:
                assert helpers.scaled_err(rec['att_val'].numpy(),
                                          g[f'layer{li}.att']) < 1e-5
            if f'layer{li}.natt' in g:
                assert helpers.scaled_err(rec['node_att_val'].numpy(),
                                          g[f'layer{li}.natt']) < 1e-5
            if f'layer{li}.m' in g:
                assert helpers.scaled_err(rec['m'].numpy(),
                                          g[f'layer{li}.m']) < 1e-5


@pytest.mark.parametrize('name', ['cfg3_k32', 'alloff_multitask',
                                  'testkwargs_fixture82'])
def test_oracle_gradients_vs_reference_golden(name):
    from oracle import egnn_oracle
    cls, kw, tasks = helpers.MODEL_GOLDENS[name]
    g, sd = helpers.load_model_golden(name)
    sd = {k: v.clone().requires_grad_(v.is_floating_point())
          for k, v in sd.items()}
    out, _ = egnn_oracle.model_forward(
        sd, torch.from_numpy(g['in.x']),
        torch.from_numpy(g['in.edge_index']).long(),
        torch.from_numpy(g['in.pos']),
        torch.from_numpy(g['in.edge_attr']).long(),
        torch.from_numpy(g['in.batch']).long(),
        num_layers=kw['num_layers'], multitask=(cls == 'multitask'),
        model_task='classification', **helpers.oracle_kwargs(kw))
    loss = torch.nn.functional.binary_cross_entropy_with_logits(
        out.reshape(-1), torch.from_numpy(g['in.y']).reshape(-1))
    loss.backward()
    assert abs(float(loss.detach()) - float(g['grad.loss'][0])) < 1e-6
    checked = 0
    for k, v in g.items():
        if not k.startswith('grad.') or k == 'grad.loss':
            continue
        got = sd[k[5:]].grad
        assert got is not None, k
        # absolute floor: some gradients are zero up to rounding noise
        err = float(np.max(np.abs(got.numpy() - v)))
        assert err < 2e-4 * float(np.max(np.abs(v))) + 2e-6, k
        checked += 1
    assert checked > 10


def test_softmax_attention_sums_to_one():
    # property of /root/reference/test/test_attention.py:22-45
    trace = []
    g, _, _ = helpers.run_oracle('testkwargs_fixture82', trace=trace)
    row = g['in.edge_index'][0]
    for rec in trace:
        sums = np.zeros(row.max() + 1)
        np.add.at(sums, row, rec['att_val'].numpy().reshape(-1))
        np.testing.assert_allclose(sums, np.ones_like(sums), atol=1e-6)
