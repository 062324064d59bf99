"""Live cross-check of the oracle against the real reference.  Runs only where
/root/reference exists (the build container); skipped on the GPU box."""
import tempfile
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(
    not ref_shim.reference_available(), reason='reference tree not present')


def test_layer_and_model_against_live_reference():
    from oracle import egnn_oracle, radius_graph as rg
    from pointvs_b200.synthetic import synthetic_complex
    ref = ref_shim.import_reference()
    torch.manual_seed(11)
    kw = dict(dim_input=13, dim_output=1, k=24, num_layers=2,
              edge_attention=True, node_attention=True, residual=True,
              normalize=True, tanh=True, graphnorm=True,
              softmax_attention=False)
    with tempfile.TemporaryDirectory() as tmp:
        model = ref.SartorrasEGNN(Path(tmp), 0, 0, None, None, silent=True,
                                  **kw).eval()
    coords, bp, feats = synthetic_complex(77, 150, 12)
    _, row, col, attr = rg.radius_graph(coords, bp, 4.0, 4.0)
    graph = ref.Data(
        x=torch.from_numpy(feats), pos=torch.from_numpy(coords).float(),
        edge_index=torch.from_numpy(np.vstack([row, col])),
        edge_attr=torch.nn.functional.one_hot(torch.from_numpy(attr).long(), 3),
        batch=torch.zeros(150, dtype=torch.long))
    pos0 = graph.pos.clone()
    with torch.no_grad():
        want = model(graph)
    # the reference mutates graph.pos in place (egnn_satorras.py:174)
    assert not torch.equal(pos0, graph.pos)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    got, x = egnn_oracle.model_forward(
        sd, graph.x, graph.edge_index, pos0, graph.edge_attr, graph.batch,
        num_layers=2, **{k: v for k, v in kw.items() if k not in (
            'dim_input', 'dim_output', 'k', 'num_layers')})
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-7)
    assert torch.allclose(x, graph.pos, rtol=0, atol=1e-6)


def test_generate_edges_against_live_reference():
    import pandas as pd
    from oracle import radius_graph as rg
    from pointvs_b200.synthetic import synthetic_complex
    ref = ref_shim.import_reference()
    coords, bp, _ = synthetic_complex(5, 500, 25)
    struct = pd.DataFrame({'x': coords[:, 0], 'y': coords[:, 1],
                           'z': coords[:, 2], 'bp': bp})
    _, (row, col), attr = ref.generate_edges(struct, 4.0, 4.0, prune=False)
    r2, c2, a2 = rg.radius_graph_c(coords, bp, 4.0, 4.0)
    np.testing.assert_array_equal(row, r2)
    np.testing.assert_array_equal(col, c2)
    np.testing.assert_array_equal(attr, a2)


def test_train_cli_flags_match_reference_parse_args(monkeypatch):
    """pointvs_b200.train.parse_args defines every flag of the reference's
    point_vs/parse_args.py with the same default."""
    import sys
    from pointvs_b200 import train
    ref_shim.install()
    from point_vs.parse_args import parse_args as ref_parse
    argv = ['multitask', '/tmp/run', '--train_data_root_pose', 'a',
            '--train_types_pose', 'b', '-ep', '2', '-k', '48', '--egnn_tanh']
    monkeypatch.setattr(sys, 'argv', ['point_vs.py'] + argv)
    want = vars(ref_parse())
    got = vars(train.parse_args(argv))
    for key, value in want.items():
        assert key in got, key
        assert got[key] == value, (key, got[key], value)
    extra = set(got) - set(want)
    assert extra == {'math', 'workers', 'worker_processes', 'host_crop'}
