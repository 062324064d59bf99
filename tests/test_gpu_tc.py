"""tcgen05 edge-MLP tiles (math='bf16x3' and 'bf16') against the oracle, the
reference goldens and the fp32 CUDA path.

bf16x3 is the error-compensated mode and must meet the north-star 1e-4
relative bound on per-complex scores; single-pass bf16 is the separately
reported fast mode (bound stated here: 3e-2 relative on scores)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from tests import helpers
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu

CFG3 = dict(dim_input=13, dim_output=1, k=64, num_layers=8,
            edge_attention=True, node_attention=True, residual=True,
            normalize=True, tanh=True, graphnorm=False)


def _layer_outputs(math, graph, seed=0, k=64, **flags):
    from pointvs_b200 import EGNNLayer
    torch.manual_seed(seed)
    layer = EGNNLayer(k, k, k, edges_in_d=3, math=math, **flags).cuda()
    with torch.no_grad():
        layer.coord_mlp[2].weight.mul_(1000.0)
    n = graph.x.shape[0]
    h = torch.randn(n, k, generator=torch.Generator().manual_seed(1)).cuda()
    coord = graph.pos.clone()
    with torch.no_grad():
        h2, x2, _, m2 = layer(h, graph.edge_index, coord, graph.edge_attr, None)
    return h2, x2, m2, layer.att_val


@pytest.mark.parametrize('k', [64, 32, 48])
def test_tc_layer_matches_fp32_layer(k):
    graph = gh.synthetic_graph(500, 3, 700, 25, ragged=True)
    flags = dict(edge_attention=True, node_attention=True, normalize=True,
                 tanh=True)
    ref = _layer_outputs('fp32', graph, k=k, **flags)
    x3 = _layer_outputs('bf16x3', graph, k=k, **flags)
    fast = _layer_outputs('bf16', graph, k=k, **flags)
    for name, a, b in zip('hxm', x3[:3], ref[:3]):
        assert helpers.scaled_err(a.cpu().numpy(), b.cpu().numpy()) < 5e-5, name
    assert helpers.scaled_err(x3[3], ref[3]) < 5e-5
    for name, a, b in zip('hxm', fast[:3], ref[:3]):
        assert helpers.scaled_err(a.cpu().numpy(), b.cpu().numpy()) < 3e-2, name


@pytest.mark.parametrize('math,tol', [('bf16x3', 1e-4), ('fp16x2', 1e-4),
                                      ('bf16', 3e-2)])
def test_tc_config3_vs_oracle(math, tol):
    model = gh.build_model(CFG3, seed=0, coord_gain=1.0)
    model.set_math(math)
    graph = gh.synthetic_graph(300, 4, 1000, 30)
    pos0 = graph.pos.clone()
    with torch.no_grad():
        out = model(graph)
    graph_cpu = SimpleNamespace(x=graph.x, pos=pos0,
                                edge_index=graph.edge_index,
                                edge_attr=graph.edge_attr, batch=graph.batch)
    want, x_want = gh.oracle_forward(model, CFG3, graph_cpu)
    err = helpers.rel_err(out.cpu().numpy().reshape(-1),
                          want.numpy().reshape(-1))
    print(f'{math}: max rel err on scores {err:.3e}')
    assert err < tol


@pytest.mark.parametrize('math', ['bf16x3', 'fp16x2'])
@pytest.mark.parametrize('name', sorted(helpers.MODEL_GOLDENS))
def test_tc_bf16x3_vs_reference_golden(name, math):
    _, _, tasks = helpers.MODEL_GOLDENS[name]
    for task in tasks:
        model, g = gh.cuda_model(name, task, math=math)
        with torch.no_grad():
            out = model(gh.cuda_graph(g))
        assert helpers.rel_err(out.cpu().numpy().reshape(-1),
                               g[f'out.{task}'].reshape(-1)) < 1e-4


def test_tc_hub_node_multi_chunk():
    from oracle import egnn_oracle
    from pointvs_b200 import EGNNLayer
    torch.manual_seed(5)
    n = 300
    layer = EGNNLayer(32, 32, 32, edges_in_d=3, edge_attention=True,
                      normalize=True, tanh=True, math='bf16x3').cuda()
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
    assert helpers.scaled_err(h2.cpu().numpy(), ho.numpy()) < 1e-4
    assert helpers.scaled_err(x2.cpu().numpy(), xo.numpy()) < 1e-4
    assert helpers.scaled_err(m2.cpu().numpy(), mo.numpy()) < 1e-4


def test_tc_deterministic():
    model = gh.build_model(CFG3, seed=0)
    model.set_math('bf16x3')
    with torch.no_grad():
        a = model(gh.synthetic_graph(300, 2, 1000, 30))
        b = model(gh.synthetic_graph(300, 2, 1000, 30))
    assert torch.equal(a, b)


@pytest.mark.parametrize('k,ragged', [(64, False), (40, True)])
def test_tc_graphnorm_node_stage_vs_oracle(k, ragged):
    """GraphNorm models run the node MLP on the tensor cores too: Wn1 -> V,
    batch statistics over all nodes, resume from V (egnn_satorras.py:84)."""
    kw = dict(dim_input=13, dim_output=1, k=k, num_layers=3, graphnorm=True,
              edge_attention=True, node_attention=True, residual=True,
              normalize=True, tanh=True)
    model = gh.build_model(kw, seed=9, coord_gain=1.0)
    with torch.no_grad():          # non-trivial affine and mean scale
        for name, p in model.named_parameters():
            if 'node_mlp.1.' in name:
                p.add_(0.3 * torch.randn_like(p))
    graph = gh.synthetic_graph(700, 5, 260, 14, ragged=ragged)
    pos0 = graph.pos.clone()
    want, x_want = gh.oracle_forward(
        model, kw, SimpleNamespace(x=graph.x, pos=pos0,
                                   edge_index=graph.edge_index,
                                   edge_attr=graph.edge_attr,
                                   batch=graph.batch))
    for math, tol in (('fp32', 1e-4), ('bf16x3', 1e-4), ('fp16x2', 1e-4),
                      ('bf16', 3e-2)):
        model.set_math(math)
        graph.pos = pos0.clone()
        with torch.no_grad():
            out = model(graph)
        assert helpers.rel_err(out.cpu().numpy().reshape(-1),
                               want.numpy().reshape(-1)) < tol, math
    # training-time recompute goes through the same kernels
    model.set_math('bf16x3')
    model.train()
    graph.pos = pos0.clone()
    out = model(graph).reshape(-1)
    out.sum().backward()
    g = model.layers[1].node_mlp[1].weight.grad
    assert g is not None and torch.isfinite(g).all() and g.abs().max() > 0


@pytest.mark.parametrize('math', ['bf16x3', 'fp16x2'])
@pytest.mark.parametrize('softmax', [False, True])
def test_tc_packed_tiles_split_nodes_and_edgeless_runs(softmax, math):
    """The tcgen05 edge kernel walks edge-packed tiles: nodes cut by a tile
    boundary (every 128 edges), hubs spanning several tiles, and runs of more
    than 128 edgeless nodes inside a tile's node range."""
    from oracle import egnn_oracle
    from pointvs_b200 import EGNNLayer
    torch.manual_seed(11)
    gen = torch.Generator().manual_seed(3)
    n = 900
    rows, cols = [], []
    # nodes 0..199: ring neighbourhoods of degree 6 (tile boundaries fall
    # inside nodes); node 200: hub with 330 edges; 201..560 edgeless; then a
    # second dense block, and an edgeless tail
    for i in range(200):
        for d in (1, 2, 3):
            rows += [i, i]
            cols += [(i + d) % 200, (i - d) % 200]
    for j in range(330):
        rows.append(200)
        cols.append(561 + j % 300)
    for i in range(561, 861):
        for d in (1, 2, 5, 7, 11):
            rows.append(i)
            cols.append(561 + (i - 561 + d) % 300)
    ei = torch.tensor([rows, cols], dtype=torch.long)
    ei = ei[:, torch.randperm(ei.shape[1], generator=gen)]   # caller order
    ea = torch.nn.functional.one_hot(
        torch.randint(0, 3, (ei.shape[1],), generator=gen), 3)
    h = torch.randn(n, 48, generator=gen)
    x = torch.randn(n, 3, generator=gen) * 4
    layer = EGNNLayer(48, 48, 48, edges_in_d=3, edge_attention=True,
                      normalize=True, tanh=True, softmax_attention=softmax,
                      math=math).cuda()
    with torch.no_grad():
        layer.coord_mlp[2].weight.mul_(300.0)
        h2, x2, _, m2 = layer(h.cuda(), ei.cuda(), x.clone().cuda(), ea.cuda())
    sd = {'l.' + k: v.detach().cpu() for k, v in layer.state_dict().items()}
    cfg = egnn_oracle.LayerConfig(residual=True, edge_attention=True,
                                  normalize=True, tanh=True,
                                  softmax_attention=softmax)
    ho, xo, mo, _ = egnn_oracle.layer_forward(sd, 'l.', cfg, h, ei[0], ei[1],
                                              x, ea)
    # per-node / per-edge outputs: fp16x2 carries the 2^-12 rounding of its
    # activation tile (stated separately from the fp32-class modes)
    tol = 5e-4 if math == 'fp16x2' else 1e-4
    assert helpers.scaled_err(h2.cpu().numpy(), ho.numpy()) < tol
    assert helpers.scaled_err(x2.cpu().numpy(), xo.numpy()) < tol
    assert helpers.scaled_err(m2.cpu().numpy(), mo.numpy()) < tol
    # edgeless nodes keep their coordinates exactly
    assert torch.equal(x2[201:561].cpu(), x[201:561])
    assert torch.equal(x2[861:].cpu(), x[861:])
