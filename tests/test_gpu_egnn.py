"""K2 (EGNN forward) on the GPU through the C ABI: parity with the reference
goldens and the CPU oracle, plus the reference's own property tests
(/root/reference/test/test_{invariance,consistency,attention}.py) re-run
against the CUDA classes.

Tolerances: per-complex scores within 1e-4 relative of the reference
(north star, fp32 mode); per-element intermediate tensors within 2e-5 of
their max magnitude."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from tests import helpers
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu

SCORE_RTOL = 1e-4
ELEM_TOL = 2e-5


@pytest.mark.parametrize('name', sorted(helpers.MODEL_GOLDENS))
def test_model_vs_reference_golden(name):
    _, _, tasks = helpers.MODEL_GOLDENS[name]
    for task in tasks:
        model, g = gh.cuda_model(name, task)
        graph = gh.cuda_graph(g)
        pos0 = graph.pos.clone()
        with torch.no_grad():
            out = model(graph)
        ref = g[f'out.{task}']
        assert tuple(out.shape) == tuple(ref.shape)
        assert helpers.rel_err(out.cpu().numpy().reshape(-1),
                               ref.reshape(-1)) < SCORE_RTOL
        n_layers = len(model.layers) - 1
        for li in range(1, n_layers + 1):
            layer = model.layers[li]
            if layer.use_coords:
                assert helpers.scaled_err(layer.intermediate_coords,
                                          g[f'layer{li}.x']) < 1e-6
            if f'layer{li}.att' in g:
                assert helpers.scaled_err(layer.att_val,
                                          g[f'layer{li}.att']) < ELEM_TOL
            if f'layer{li}.natt' in g:
                assert helpers.scaled_err(layer.node_att_val,
                                          g[f'layer{li}.natt']) < ELEM_TOL
        # in-place coordinate update of an fp32 on-device pos (reference:
        # egnn_satorras.py:174 through pnn_geometric_base.py:56-57)
        if model.layers[1].use_coords:
            assert not torch.equal(pos0, graph.pos)
            assert helpers.scaled_err(graph.pos.cpu().numpy(),
                                      g[f'layer{n_layers}.x']) < 1e-6
        else:
            assert torch.equal(pos0, graph.pos)


@pytest.mark.parametrize('name', ['cfg3_k32', 'gated_tanhatt_multifc',
                                  'rezero_perminv_static',
                                  'testkwargs_fixture82'])
def test_get_embeddings_vs_reference_golden(name):
    """h_L and the final messages m_L (caller's edge order)."""
    model, g = gh.cuda_model(name)
    graph = gh.cuda_graph(g)
    with torch.no_grad():
        h, m = model.get_embeddings(graph.x, graph.edge_index, graph.pos,
                                    graph.edge_attr, graph.batch)
    n_layers = len(model.layers) - 1
    assert helpers.scaled_err(h.cpu().numpy(),
                              g[f'layer{n_layers}.h']) < ELEM_TOL
    assert helpers.scaled_err(m.cpu().numpy(),
                              g[f'layer{n_layers}.m']) < ELEM_TOL


def test_layer_api_config2():
    """BASELINE config 2: one EGNNLayer(32,32,32, edges_in_d=3), 16 x 400
    atoms, h ~ N(0,1): h', x', m against the oracle on the same inputs."""
    from oracle import egnn_oracle
    from pointvs_b200 import EGNNLayer
    torch.manual_seed(0)
    layer = EGNNLayer(32, 32, 32, edges_in_d=3, edge_attention=True,
                      node_attention=True, normalize=True, tanh=True).cuda()
    with torch.no_grad():
        layer.coord_mlp[2].weight.mul_(1000.0)
    graph = gh.synthetic_graph(200, 16, 400, 25)
    n = graph.x.shape[0]
    h = torch.randn(n, 32, generator=torch.Generator().manual_seed(1)).cuda()
    ei, ea = graph.edge_index, graph.edge_attr
    assert 90_000 < ei.shape[1] < 105_000
    coord = graph.pos.clone()
    with torch.no_grad():
        h2, x2, ea2, m2 = layer(h, ei, coord, ea, None)
    assert x2 is coord          # in place, like the reference
    sd = {'l.' + k: v.detach().cpu() for k, v in layer.state_dict().items()}
    cfg = egnn_oracle.LayerConfig(residual=True, edge_attention=True,
                                  node_attention=True, normalize=True,
                                  tanh=True)
    ho, xo, mo, side = egnn_oracle.layer_forward(
        sd, 'l.', cfg, h.cpu(), ei[0].cpu(), ei[1].cpu(), graph.pos.cpu(),
        ea.cpu())
    assert helpers.scaled_err(h2.cpu().numpy(), ho.numpy()) < ELEM_TOL
    assert helpers.scaled_err(m2.cpu().numpy(), mo.numpy()) < ELEM_TOL
    assert helpers.scaled_err(x2.cpu().numpy(), xo.numpy()) < 1e-6
    assert float((xo - graph.pos.cpu()).abs().max()) > 1e-3   # coords do move
    assert helpers.scaled_err(layer.att_val,
                              side['att_val'].numpy()) < ELEM_TOL


@pytest.mark.parametrize('ragged', [False, True])
def test_config3_vs_oracle(ragged):
    """BASELINE config 3 model (8 layers, k 64, edge+node attention,
    residual, normalise, tanh) on 1000-atom complexes, coordinate head at
    gain 1 so the coordinate path matters."""
    kw = dict(dim_input=13, dim_output=1, k=64, num_layers=8,
              edge_attention=True, node_attention=True, residual=True,
              normalize=True, tanh=True, graphnorm=False)
    model = gh.build_model(kw, seed=0, coord_gain=1.0)
    graph = gh.synthetic_graph(300, 4, 1000, 30, ragged=ragged)
    pos0 = graph.pos.clone()
    with torch.no_grad():
        out = model(graph)
    graph_cpu = SimpleNamespace(x=graph.x, pos=pos0,
                                edge_index=graph.edge_index,
                                edge_attr=graph.edge_attr, batch=graph.batch)
    want, x_want = gh.oracle_forward(model, kw, graph_cpu)
    assert helpers.rel_err(out.cpu().numpy().reshape(-1),
                           want.numpy().reshape(-1)) < SCORE_RTOL
    assert helpers.scaled_err(graph.pos.cpu().numpy(), x_want.numpy()) < 1e-5


def test_alloff_config1_vs_oracle():
    """Config-1 style: 3 layers, k 32, everything switched off, dim_input 22,
    estimate_bonds radii (degree-0 nodes exercise the clamped mean)."""
    kw = dict(dim_input=13, dim_output=1, k=32, num_layers=3,
              edge_attention=False, node_attention=False, residual=False,
              normalize=False, tanh=False, graphnorm=False)
    model = gh.build_model(kw, multitask=True, seed=3, coord_gain=1.0)
    graph = gh.synthetic_graph(400, 8, 300, 20, radii=(4.0, 2.0), ragged=True)
    deg = np.diff(graph.pvs_csr.row_ptr.cpu().numpy())
    assert (deg == 0).any()
    pos0 = graph.pos.clone()
    with torch.no_grad():
        out = model(graph)
    graph_cpu = SimpleNamespace(x=graph.x, pos=pos0,
                                edge_index=graph.edge_index,
                                edge_attr=graph.edge_attr, batch=graph.batch)
    want, _ = gh.oracle_forward(model, kw, graph_cpu, multitask=True)
    assert helpers.rel_err(out.cpu().numpy().reshape(-1),
                           want.numpy().reshape(-1)) < SCORE_RTOL


def test_edge_order_independence_and_side_channel_order():
    """Attribution passes arbitrarily ordered edge_index
    (attribution_fns.py:80-101); scores must not depend on the order and
    att_val must come back in the caller's order."""
    model, g = gh.cuda_model('cfg3_k32')
    graph = gh.cuda_graph(g)
    with torch.no_grad():
        base = model(graph).cpu().numpy()
    att_base = model.layers[1].att_val
    perm = torch.randperm(graph.edge_index.shape[1],
                          generator=torch.Generator().manual_seed(3)).cuda()
    graph2 = gh.cuda_graph(g)
    graph2.edge_index = graph2.edge_index[:, perm].contiguous()
    graph2.edge_attr = graph2.edge_attr[perm].contiguous()
    with torch.no_grad():
        out = model(graph2).cpu().numpy()
    assert helpers.rel_err(out, base) < 1e-5
    np.testing.assert_allclose(model.layers[1].att_val,
                               att_base[perm.cpu().numpy()], atol=1e-6)


def _rotation(seed):
    q, _ = np.linalg.qr(np.random.default_rng(seed).normal(size=(3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return torch.from_numpy(q.astype(np.float32))


def test_e3_invariance():
    """test/test_invariance.py:35-43 on the CUDA class (tolerance 3e-5 on the
    sigmoid), extended with translation and reflection."""
    model, g = gh.cuda_model('testkwargs_fixture82')
    graph = gh.cuda_graph(g)
    with torch.no_grad():
        base = torch.sigmoid(model(graph)).cpu().numpy()
    rot = _rotation(2).cuda()
    refl = torch.diag(torch.tensor([1.0, 1.0, -1.0])).cuda()
    for transform in (lambda p: p @ rot,
                      lambda p: p + torch.tensor([3.0, -2.0, 5.0]).cuda(),
                      lambda p: p @ refl,
                      lambda p: (p @ rot) @ refl + 1.5):
        graph2 = gh.cuda_graph(g)
        graph2.pos = transform(graph2.pos).contiguous()
        with torch.no_grad():
            got = torch.sigmoid(model(graph2)).cpu().numpy()
        np.testing.assert_allclose(got, base, atol=3e-5)


def test_consistency():
    """test/test_consistency.py:23-33: repeated forwards agree (pos is
    refreshed each time; the reference test relies on gain 0.001 instead)."""
    model, g = gh.cuda_model('testkwargs_fixture82')
    with torch.no_grad():
        base = float(torch.sigmoid(model(gh.cuda_graph(g)))[0])
    assert abs(base) > 1e-5
    for _ in range(10):
        with torch.no_grad():
            got = float(torch.sigmoid(model(gh.cuda_graph(g)))[0])
        assert got == pytest.approx(base, abs=3e-5)
    # bitwise: the kernels use no atomics on floating-point data
    with torch.no_grad():
        a = model(gh.cuda_graph(g))
        b = model(gh.cuda_graph(g))
    assert torch.equal(a, b)


def test_softmax_attention_sums_to_one():
    """test/test_attention.py:22-45 on the CUDA class, batch of two graphs."""
    model, g = gh.cuda_model('testkwargs_fixture82')
    graph = gh.cuda_graph(g)
    with torch.no_grad():
        model(graph)
    row = g['in.edge_index'][0]
    checked = False
    for layer in model.layers:
        if hasattr(layer, 'att_val') and layer.att_val is not None:
            checked = True
            sums = np.zeros(row.max() + 1)
            np.add.at(sums, row, layer.att_val.squeeze())
            np.testing.assert_allclose(sums, np.ones_like(sums), atol=1e-6)
    assert checked


def test_single_graph_output_shape():
    """B = 1 returns [dim_output] (pnn_geometric_base.py:31-32)."""
    kw = dict(dim_input=13, dim_output=1, k=32, num_layers=2,
              graphnorm=False)
    model = gh.build_model(kw)
    graph = gh.synthetic_graph(1, 1, 200, 10)
    with torch.no_grad():
        out = model(graph)
    assert tuple(out.shape) == (1,)
    model3 = gh.build_model(dict(kw, dim_output=3))
    with torch.no_grad():
        assert tuple(model3(gh.synthetic_graph(1, 1, 200, 10)).shape) == (3,)
        assert tuple(model3(gh.synthetic_graph(1, 2, 200, 10)).shape) == (2, 3)


def test_cpu_tensors_fail_loudly():
    from pointvs_b200._cabi import PvsError
    from pointvs_b200 import EGNNLayer
    layer = EGNNLayer(32, 32, 32, edges_in_d=3)
    h = torch.zeros(4, 32)
    ei = torch.tensor([[0, 1], [1, 0]])
    with pytest.raises(PvsError):
        layer(h, ei, torch.zeros(4, 3), torch.zeros(2, 3, dtype=torch.long))


def test_unsupported_width_fails_loudly():
    from pointvs_b200._cabi import PvsError
    kw = dict(dim_input=13, dim_output=1, k=96, num_layers=1, graphnorm=False)
    model = gh.build_model(kw)
    with pytest.raises(PvsError):
        with torch.no_grad():
            model(gh.synthetic_graph(1, 1, 100, 10))


def test_high_degree_node_multi_chunk():
    """A hub with > 128 incoming edges spans several 128-edge chunks."""
    from oracle import egnn_oracle
    from pointvs_b200 import EGNNLayer
    torch.manual_seed(5)
    n = 300
    layer = EGNNLayer(32, 32, 32, edges_in_d=3, edge_attention=True,
                      normalize=True, tanh=True).cuda()
    with torch.no_grad():
        layer.coord_mlp[2].weight.mul_(1000.0)
    hub = torch.zeros(n - 1, dtype=torch.long)
    others = torch.arange(1, n)
    ei = torch.cat([torch.stack([hub, others]), torch.stack([others, hub])], 1)
    gen = torch.Generator().manual_seed(2)
    ea = torch.nn.functional.one_hot(
        torch.randint(0, 3, (ei.shape[1],), generator=gen), 3)
    h = torch.randn(n, 32, generator=gen)
    x = torch.randn(n, 3, generator=gen) * 3
    with torch.no_grad():
        h2, x2, _, m2 = layer(h.cuda(), ei.cuda(), x.clone().cuda(), ea.cuda())
    sd = {'l.' + k: v.detach().cpu() for k, v in layer.state_dict().items()}
    cfg = egnn_oracle.LayerConfig(residual=True, edge_attention=True,
                                  normalize=True, tanh=True)
    ho, xo, mo, _ = egnn_oracle.layer_forward(sd, 'l.', cfg, h, ei[0], ei[1],
                                              x, ea)
    assert helpers.scaled_err(h2.cpu().numpy(), ho.numpy()) < ELEM_TOL
    assert helpers.scaled_err(x2.cpu().numpy(), xo.numpy()) < 1e-5
    assert helpers.scaled_err(m2.cpu().numpy(), mo.numpy()) < ELEM_TOL


def test_capacity_bounded_batch_scores_identically():
    from pointvs_b200.graph import PackedBatch
    from pointvs_b200.synthetic import synthetic_batch
    kw = dict(dim_input=13, dim_output=1, k=64, num_layers=3,
              edge_attention=True, node_attention=True, residual=True,
              normalize=True, tanh=True, graphnorm=False)
    model = gh.build_model(kw, seed=1, coord_gain=1.0)
    coords, bp, feats, cptr = synthetic_batch(70, 3, 500, 20, ragged=True)
    outs = []
    for cap in (None, 'auto'):
        batch = PackedBatch.from_arrays(coords, bp, feats, cptr, 4.0, 4.0,
                                        edge_capacity=cap)
        with torch.no_grad():
            outs.append(model(batch))
        batch.pvs_csr.check_overflow()
    assert torch.equal(outs[0], outs[1])


def test_config5_pocket_poses_vs_oracle():
    """BASELINE configs[4] shape: one 800-atom pocket, many 30-atom ligand
    poses (830 atoms per complex), scored in one packed batch."""
    from pointvs_b200.graph import PackedBatch
    from pointvs_b200.synthetic import synthetic_pocket_poses
    kw = dict(dim_input=13, dim_output=1, k=64, num_layers=8,
              edge_attention=True, node_attention=True, residual=True,
              normalize=True, tanh=True, graphnorm=False)
    model = gh.build_model(kw, seed=0, coord_gain=1.0)
    coords, bp, feats, cptr = synthetic_pocket_poses(0, 6)
    assert cptr[1] == 830
    batch = PackedBatch.from_arrays(coords, bp, feats, cptr, 4.0, 4.0)
    pos0 = batch.pos.clone()
    with torch.no_grad():
        out = model(batch)
    assert 10.0 < batch.pvs_csr.n_edges / (6 * 830) < 20.0
    graph_cpu = SimpleNamespace(x=batch.x, pos=pos0,
                                edge_index=batch.edge_index,
                                edge_attr=batch.edge_attr, batch=batch.batch)
    want, _ = gh.oracle_forward(model, kw, graph_cpu)
    assert helpers.rel_err(out.cpu().numpy().reshape(-1),
                           want.numpy().reshape(-1)) < SCORE_RTOL
    # poses differ, so scores must differ
    assert float(out.std()) > 0


@pytest.mark.parametrize('name', ['cfg3_k32', 'gated_tanhatt_multifc',
                                  'multitask_firstfinal', 'alloff_multitask',
                                  'testkwargs_fixture82',
                                  'rezero_perminv_static'])
def test_one_call_scoring_path_equals_layerwise_path(name):
    """pvs_egnn_model_fwd (side channels off) vs the per-layer calls: same
    kernels, so bitwise equal scores and the same in-place coordinate update;
    and still within 1e-4 of the reference golden."""
    _, _, tasks = helpers.MODEL_GOLDENS[name]
    for task in tasks:
        model, g = gh.cuda_model(name, task)
        slow_graph = gh.cuda_graph(g)
        with torch.no_grad():
            slow = model(slow_graph)
            assert not model._fast_path_ok()     # side channels are on
        model.set_record_side_channels(False)
        model.record_embed_coords = False
        fast_graph = gh.cuda_graph(g)
        with torch.no_grad():
            assert model._fast_path_ok()
            fast = model(fast_graph)
        assert torch.equal(fast, slow)
        assert torch.equal(fast_graph.pos, slow_graph.pos)
        assert helpers.rel_err(fast.cpu().numpy().reshape(-1),
                               g[f'out.{task}'].reshape(-1)) < SCORE_RTOL
        # with grad enabled the autograd path must be taken
        assert not model._fast_path_ok()


def test_one_call_path_sees_in_place_parameter_updates():
    kw = dict(dim_input=13, dim_output=1, k=32, num_layers=2, graphnorm=False,
              edge_attention=True)
    model = gh.build_model(kw, seed=4)
    model.set_record_side_channels(False)
    model.record_embed_coords = False
    with torch.no_grad():
        a = model(gh.synthetic_graph(3, 2, 150, 10))
        for p in model.parameters():
            p.mul_(1.1)                     # what an optimiser step does
        b = model(gh.synthetic_graph(3, 2, 150, 10))
        model.set_record_side_channels(True)
        c = model(gh.synthetic_graph(3, 2, 150, 10))
    assert not torch.equal(a, b)
    assert torch.equal(b, c)


@pytest.mark.parametrize('math', ['fp32', 'bf16x3'])
def test_edgeless_and_tiny_graphs(math):
    """Graphs with no edges at all, single atoms, and fewer nodes than a tile."""
    from pointvs_b200.graph import PackedBatch
    kw = dict(dim_input=13, dim_output=1, k=64, num_layers=2,
              edge_attention=True, node_attention=True, residual=True,
              normalize=True, tanh=True, graphnorm=False)
    model = gh.build_model(kw, seed=2, coord_gain=1.0)
    model.set_math(math)
    rng = np.random.default_rng(0)
    # complex 0: three atoms 50 A apart (no edges); complex 1: one atom;
    # complex 2: a 5-atom cluster
    coords = np.concatenate([np.eye(3) * 50.0, np.zeros((1, 3)) + 200.0,
                             rng.normal(size=(5, 3)) + 400.0])
    coords = coords.astype(np.float32).astype(np.float64)
    bp = np.array([0, 1, 1, 1, 0, 0, 1, 1, 1])
    feats = np.zeros((9, 13), dtype=np.float32)
    feats[np.arange(9), rng.integers(0, 12, 9)] = 1.0
    batch = PackedBatch.from_arrays(coords, bp, feats, [0, 3, 4, 9], 4.0, 4.0)
    pos0 = batch.pos.clone()
    with torch.no_grad():
        out = model(batch)
    assert out.shape == (3, 1) and bool(torch.isfinite(out).all())
    graph_cpu = SimpleNamespace(x=batch.x, pos=pos0, edge_index=batch.edge_index,
                                edge_attr=batch.edge_attr, batch=batch.batch)
    want, x_want = gh.oracle_forward(model, kw, graph_cpu)
    assert helpers.rel_err(out.cpu().numpy().reshape(-1),
                           want.numpy().reshape(-1)) < SCORE_RTOL
    # atoms without neighbours do not move
    assert torch.equal(batch.pos[:4], pos0[:4])
    assert helpers.scaled_err(batch.pos.cpu().numpy(), x_want.numpy()) < 1e-5


def test_val_writes_reference_prediction_file(tmp_path):
    """PointNeuralNetworkBase.val on a list of packed batches: file name and
    line format of the reference (point_neural_network_base.py:208-325)."""
    import pointvs_b200 as pv
    from pointvs_b200.synthetic import synthetic_batch
    kw = dict(dim_input=13, dim_output=1, k=32, num_layers=2, graphnorm=False)
    torch.manual_seed(0)
    model = pv.SartorrasEGNN(tmp_path, 0, 0, None, None, **kw).cuda()
    loader = []
    for b in range(3):
        coords, bp, feats, cptr = synthetic_batch(10 * b, 2, 120, 10)
        batch = pv.PackedBatch.from_arrays(
            coords, bp, feats, cptr, 4.0, 4.0,
            y=torch.tensor([1.0, 0.0]),
            lig_fname=[f'lig_{b}_0.parquet', f'lig_{b}_1.parquet'],
            rec_fname=[f'rec_{b}.parquet'] * 2)
        loader.append(batch)
    model.val(loader, tmp_path / 'predictions_test.txt')
    lines = (tmp_path / 'pose_predictions_test.txt').read_text().splitlines()
    assert len(lines) == 6
    for i, line in enumerate(lines):
        y_true, rest = line.split(' | ')
        y_pred, rec, lig = rest.split(' ')
        assert y_true == ('1.000' if i % 2 == 0 else '0.000')
        assert 0.0 <= float(y_pred) <= 1.0 and len(y_pred.split('.')[1]) == 3
        assert rec == f'rec_{i // 2}.parquet' and lig == f'lig_{i // 2}_{i % 2}.parquet'


def test_edge_dropout_matches_published_dropout_adj_semantics():
    """Training-time edge dropout (egnn_satorras.py:320-323, PyG 2.0.4
    dropout_adj with force_undirected=True): survivors are a symmetric subset
    of the original edges, about (1 - p) of the undirected pairs; p = 0 and
    eval mode leave the graph alone; a forward on the thinned graph equals the
    oracle on that edge list."""
    from pointvs_b200.graph import dropout_adj
    kw = dict(dim_input=13, dim_output=1, k=32, num_layers=2, graphnorm=False,
              edge_attention=True, node_attention=True, residual=True,
              normalize=True, tanh=True, dropout=0.4)
    model = gh.build_model(kw, seed=3, coord_gain=1.0)
    graph = gh.synthetic_graph(400, 3, 300, 15)
    csr = graph.pvs_csr
    base = set(zip(*[t.tolist() for t in csr.edge_index('csr').cpu()],
                   csr.attr.cpu().tolist()))
    gen = torch.Generator(device='cuda').manual_seed(7)
    thin = dropout_adj(csr, 0.4, generator=gen)
    ei = thin.edge_index('csr').cpu()
    got = set(zip(ei[0].tolist(), ei[1].tolist(), thin.attr.cpu().tolist()))
    assert got <= base
    assert all((c, r, a) in got for r, c, a in got)          # mirrored
    und = sum(1 for r, c, _ in base if r < c)
    frac = (thin.n_edges / 2) / und
    assert abs(frac - 0.6) < 0.03
    same = dropout_adj(csr, 0.0)
    assert same.n_edges == 2 * und          # every undirected pair, mirrored
    # eval mode: dropout is off, scores equal the dropout-free model's
    pos0 = graph.pos.clone()
    with torch.no_grad():
        a = model(graph)
    model.dropout_p = 0.0
    graph.pos = pos0.clone()
    with torch.no_grad():
        b = model(graph)
    assert torch.equal(a, b)
    # training mode: the layers run on the thinned graph
    model.dropout_p = 0.4
    model.train()
    torch.manual_seed(11)
    graph.pos = pos0.clone()
    with torch.no_grad():
        h_got, _ = model.get_embeddings(graph.x, None, graph.pos, None,
                                        graph.batch, _csr=csr,
                                        _want_messages=False)
    torch.manual_seed(11)
    thin = dropout_adj(csr, 0.4)
    model.eval()
    graph.pos = pos0.clone()
    with torch.no_grad():
        h_want, _ = model.get_embeddings(graph.x, None, graph.pos, None,
                                         graph.batch, _csr=thin,
                                         _want_messages=False)
    assert torch.equal(h_got, h_want)
    from oracle import egnn_oracle
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    cfgs = egnn_oracle.layer_configs(2, **helpers.oracle_kwargs(kw))
    h_ref, _, _ = egnn_oracle.embeddings(
        sd, cfgs, graph.x.cpu(), thin.edge_index('csr').cpu(), pos0.cpu(),
        thin.edge_attr_onehot('csr').cpu())
    assert helpers.scaled_err(h_want.cpu().numpy(), h_ref.numpy()) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize('ki,ko,act', [(13, 64, 'none'), (20, 48, 'silu'),
                                       (32, 64, 'none'), (7, 30, 'sigmoid')])
def test_streaming_small_k_linear_equals_tiled_linear(ki, ko, act):
    """The streaming form of the linear (short inputs, >= 4096 rows: the
    atom-feature embedding) against fp64 PyTorch (1e-6) and, bit for bit,
    against the tiled kernel that smaller row counts take (same accumulation
    order); and the vectorised mean pool against PyTorch."""
    import torch
    from pointvs_b200 import dense
    gen = torch.Generator().manual_seed(ki * 100 + ko)
    x = torch.randn(5003, ki, generator=gen).cuda()
    w = (torch.randn(ko, ki, generator=gen) * 0.3).cuda()
    b = torch.randn(ko, generator=gen).cuda()
    big = dense.linear(x, w, b, act)                       # streaming kernel
    small = torch.cat([dense.linear(x[i:i + 1000], w, b, act)    # tiled kernel
                       for i in range(0, 5003, 1000)])
    assert torch.equal(big, small)
    ref = x.double() @ w.double().t() + b.double()
    ref = {'none': ref, 'silu': torch.nn.functional.silu(ref),
           'sigmoid': torch.sigmoid(ref)}[act]
    assert float((big.double() - ref).abs().max()) < 2e-6 * max(1.0, float(ref.abs().max()))


@pytest.mark.gpu
def test_mean_pool_matches_torch():
    import torch
    from pointvs_b200 import dense
    gen = torch.Generator().manual_seed(3)
    sizes = [1, 17, 1000, 33, 512, 5]
    ptr_ = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int32).cuda()
    for k in (64, 40):
        h = torch.randn(int(ptr_[-1]), k, generator=gen).cuda()
        got = dense.mean_pool(h, None, len(sizes), graph_ptr=ptr_)
        want = torch.stack([h[int(ptr_[i]):int(ptr_[i + 1])].double().mean(0)
                            for i in range(len(sizes))])
        assert float((got.double() - want).abs().max()) < 1e-6
