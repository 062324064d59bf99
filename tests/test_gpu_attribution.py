"""Row N3: attribution callers (pointvs_b200/attribution.py) against arrays
returned by the reference's own attribution_fns on the same complex and
weights (tests/golden/attribution.npz), and the coordinate-tracking ones
against the oracle's per-layer trace."""
from pathlib import Path

import numpy as np
import pytest
import torch

from tests import helpers
from tests.golden.make_attribution_cfg import MODELS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def gold():
    return helpers.load_npz('attribution.npz')


def _model(gold, name):
    import pointvs_b200 as pv
    kw = MODELS[name]
    model = pv.SartorrasEGNN(Path('/tmp/pvs_test'), 0, 0, None, None,
                             silent=True, **kw)
    sd = {k[len(name) + 4:]: torch.from_numpy(v) for k, v in gold.items()
          if k.startswith(f'{name}/sd.')}
    model.load_state_dict(sd)
    return model.cuda().eval(), sd, kw


def _inputs(gold):
    return dict(
        p=torch.from_numpy(gold['pos']).unsqueeze(0).cuda(),
        v=torch.from_numpy(gold['x']).unsqueeze(0).cuda(),
        edge_indices=torch.from_numpy(gold['edge_index'].astype(np.int64)).cuda(),
        edge_attrs=torch.from_numpy(gold['edge_attr'].astype(np.int64)).cuda())


@pytest.mark.parametrize('name,sig,bs', [('plain', False, 32),
                                         ('plain', True, 173),
                                         ('plain', False, 7),
                                         ('gn1', False, 32),
                                         ('gn1', True, 32)])
def test_atom_masking_vs_reference(gold, name, sig, bs):
    from pointvs_b200 import attribution as A
    model, _, _ = _model(gold, name)
    got = A.atom_masking(model, bs=bs, sigmoid=sig, **_inputs(gold))
    want = gold[f'{name}/atom_masking/sigmoid{int(sig)}']
    assert got.shape == want.shape
    # differences of O(1) scores: fp32 round-off of the scores themselves
    np.testing.assert_allclose(got, want, atol=2e-6, rtol=2e-3)


def test_bond_masking_vs_reference(gold):
    from pointvs_b200 import attribution as A
    model, _, _ = _model(gold, 'gn3')
    got = A.bond_masking(model, bs=16, **_inputs(gold))
    want = gold['gn3/bond_masking/sigmoid0']
    assert got.shape == want.shape
    inter = gold['edge_attr'][:, 1] != 0
    assert np.all(got[~inter] == 0) and np.all(want[~inter] == 0)
    np.testing.assert_allclose(got, want, atol=2e-6, rtol=2e-3)


def test_bond_masking_single_output_and_batching_consistency(gold):
    """The reference raises for single-output models; here it works, and the
    batch size must not change any score."""
    from pointvs_b200 import attribution as A
    model, _, _ = _model(gold, 'plain')
    a = A.bond_masking(model, bs=64, **_inputs(gold))
    b = A.bond_masking(model, bs=5, **_inputs(gold))
    np.testing.assert_allclose(a, b, atol=1e-6)
    assert np.abs(a).max() > 0


@pytest.mark.parametrize('name', ['plain', 'gn3', 'gn1'])
def test_attention_attributions_vs_reference(gold, name):
    from pointvs_b200 import attribution as A
    model, _, kw = _model(gold, name)
    got = A.edge_attention(model, **_inputs(gold))
    np.testing.assert_allclose(got, gold[f'{name}/edge_attention/sigmoid0'],
                               atol=2e-6, rtol=1e-5)
    ranks = A.mean_edge_attention_rank(model, **_inputs(gold))
    d = np.abs(ranks - gold[f'{name}/mean_edge_attention_rank/sigmoid0'])
    # ranks of near-tied weights may swap under fp32 round-off
    assert d.max() <= 0.02 * len(d) and d.mean() < 0.5
    assert np.corrcoef(
        ranks, gold[f'{name}/mean_edge_attention_rank/sigmoid0'])[0, 1] > 0.9999
    if kw['node_attention']:
        got = A.node_attention(model, **_inputs(gold))
        np.testing.assert_allclose(
            got, gold[f'{name}/node_attention/sigmoid0'], atol=2e-6, rtol=1e-5)
        got = A.node_attention(model, sigmoid=True, **_inputs(gold))
        np.testing.assert_allclose(
            got, gold[f'{name}/node_attention/sigmoid1'], atol=2e-5, rtol=1e-4)
        ranks = A.mean_node_attention_rank(model, **_inputs(gold))
        d = np.abs(ranks - gold[f'{name}/mean_node_attention_rank/sigmoid0'])
        assert d.max() <= 0.05 * len(d) and d.mean() < 0.5
    got = A.cam(model, **_inputs(gold))
    want = gold[f'{name}/cam/sigmoid0']
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, atol=5e-6, rtol=1e-4)


def test_coordinate_tracking_vs_oracle_trace(gold):
    from oracle import egnn_oracle
    from pointvs_b200 import attribution as A
    model, sd, kw = _model(gold, 'plain')
    trace = []
    ei = torch.from_numpy(gold['edge_index'].astype(np.int64))
    egnn_oracle.model_forward(
        sd, torch.from_numpy(gold['x']), ei, torch.from_numpy(gold['pos']),
        torch.from_numpy(gold['edge_attr'].astype(np.int64)),
        torch.zeros(len(gold['x']), dtype=torch.long),
        num_layers=kw['num_layers'], trace=trace, **helpers.oracle_kwargs(kw))
    x0 = gold['pos']
    want = sum(np.sqrt(((t['x'].numpy() - x0) ** 2).sum(1)) for t in trace)
    assert want.max() > 1e-3           # the coordinate path is visible
    got = A.track_position_changes(model, **_inputs(gold))
    np.testing.assert_allclose(got, want, atol=1e-5, rtol=1e-4)
    xl = trace[-1]['x'].numpy()
    e = ei.numpy()
    want = np.linalg.norm(xl[e[0]] - xl[e[1]], axis=1) - \
        np.linalg.norm(x0[e[0]] - x0[e[1]], axis=1)
    got = A.track_bond_lengths(model, **_inputs(gold))
    np.testing.assert_allclose(got, want, atol=2e-5)


def test_masked_copies_layout():
    from pointvs_b200.attribution import masked_copies
    x = torch.arange(5, dtype=torch.float32).unsqueeze(1).cuda()
    pos = torch.zeros(5, 3).cuda()
    ei = torch.tensor([[0, 1, 1, 2, 3, 4], [1, 0, 2, 1, 4, 3]]).cuda()
    ea = torch.eye(3, dtype=torch.long)[[0, 0, 1, 1, 2, 2]].cuda()
    removed = torch.tensor([[1, -1], [0, 4]]).cuda()
    g = masked_copies(x, pos, ei, ea, removed)
    assert g.x.reshape(-1).tolist() == [0, 2, 3, 4, 1, 2, 3]
    assert g.batch.tolist() == [0, 0, 0, 0, 1, 1, 1]
    # copy 0 keeps the 3-4 edges (renumbered 2-3); copy 1 keeps 1-2 (4-5)
    assert g.edge_index.tolist() == [[2, 3, 4, 5], [3, 2, 5, 4]]
    assert g.edge_attr.argmax(1).tolist() == [2, 2, 1, 1]
