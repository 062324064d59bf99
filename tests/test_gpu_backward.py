"""K3 (backward) on the GPU through the C ABI: parameter and input gradients
against the reference's own autograd (golden fixtures) and against autograd of
the CPU oracle on seeded inputs.

Tolerance: max |g - g_ref| <= 2e-4 * max|g_ref| + 2e-6 per parameter tensor
(fp32 sums over ~1e5 edges in a different order than the reference)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from tests import helpers
from tests import gpu_helpers as gh

pytestmark = pytest.mark.gpu


def _close(got, want, name, rtol=2e-4, atol=2e-6):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, name
    err = float(np.max(np.abs(got - want)))
    bound = rtol * float(np.max(np.abs(want))) + atol
    assert err <= bound, f'{name}: err {err:.3e} > {bound:.3e}'


@pytest.mark.parametrize('math', ['fp32', 'bf16x3'])
@pytest.mark.parametrize('name', ['cfg3_k32', 'alloff_multitask',
                                  'testkwargs_fixture82'])
def test_param_grads_vs_reference_golden(name, math):
    """math = bf16x3: forward, recompute and the edge backward run on tcgen05
    (egnn_edge_bwd_tc.cu; configurations it does not cover -- softmax
    attention in `testkwargs_fixture82` -- stay on the FFMA edge backward)."""
    model, g = gh.cuda_model(name, 'classification', math=math)
    model.train()
    graph = gh.cuda_graph(g)
    out = model(graph).reshape(-1)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(
        out, graph.y.reshape(-1))
    loss.backward()
    assert abs(float(loss) - float(g['grad.loss'][0])) < 1e-5
    checked = 0
    for pname, p in model.named_parameters():
        key = 'grad.' + pname
        if key not in g:
            continue
        assert p.grad is not None, pname
        _close(p.grad.cpu().numpy(), g[key], pname)
        checked += 1
    assert checked > 10


VARIANTS = {
    'cfg3_k64': dict(k=64, num_layers=3, edge_attention=True,
                     node_attention=True, residual=True, normalize=True,
                     tanh=True),
    'k48_relu_att_noresid': dict(k=48, num_layers=2, edge_attention=True,
                                 node_attention=True, residual=False,
                                 normalize=False, tanh=False,
                                 attention_activation_fn='relu'),
    'edge_residual_plain': dict(k=32, num_layers=3, edge_attention=True,
                                node_attention=False, residual=True,
                                edge_residual=True, normalize=True, tanh=True,
                                attention_activation_fn='tanh'),
    'rezero': dict(k=32, num_layers=3, edge_attention=False,
                   node_attention=True, residual=True, edge_residual=True,
                   rezero=True, normalize=True, tanh=True),
    'gated': dict(k=32, num_layers=3, edge_attention=True, node_attention=True,
                  residual=True, edge_residual=True, gated_residual=True,
                  normalize=True, tanh=False,
                  attention_activation_fn='silu'),
    'graphnorm_softmax': dict(k=32, num_layers=3, edge_attention=True,
                              node_attention=True, residual=True,
                              normalize=True, tanh=True, graphnorm=True,
                              softmax_attention=True),
    'graphnorm_only': dict(k=64, num_layers=2, edge_attention=True,
                           node_attention=False, residual=True,
                           normalize=True, tanh=True, graphnorm=True),
    'softmax_only_eres': dict(k=32, num_layers=3, edge_attention=True,
                              node_attention=True, residual=True,
                              edge_residual=True, normalize=False, tanh=True,
                              softmax_attention=True),
    'perminv_static': dict(k=32, num_layers=2, edge_attention=True,
                           node_attention=False, residual=True,
                           normalize=True, tanh=True,
                           permutation_invariance=True, update_coords=False),
}


@pytest.mark.parametrize('math', ['fp32', 'bf16x3'])
@pytest.mark.parametrize('vname', sorted(VARIANTS))
def test_grads_vs_oracle_autograd(vname, math):
    kw = dict(dim_input=13, dim_output=1, **{'graphnorm': False, **VARIANTS[vname]})
    model = gh.build_model(kw, seed=7, coord_gain=1.0)
    model.set_math(math)
    gen = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for pname, p in model.named_parameters():
            if 'gate_parameter' in pname:
                p.copy_(torch.rand(1, generator=gen).cuda() * 0.8 + 0.1)
            if '.node_mlp.1.' in pname:     # GraphNorm weight / bias / mean_scale
                p.copy_((torch.rand(p.shape, generator=gen) + 0.5).cuda())
    model.train()
    radii = (4.0, 2.0) if vname == 'k48_relu_att_noresid' else (4.0, 4.0)
    graph = gh.synthetic_graph(900, 3, 250, 15, radii=radii, ragged=True)
    y = torch.tensor([1.0, 0.0, 1.0], device='cuda')
    pos0 = graph.pos.clone()
    graph.pos = graph.pos.clone().requires_grad_(True)
    out = model(graph).reshape(-1)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y)
    loss.backward()

    # oracle autograd on the CPU with the same weights
    from oracle import egnn_oracle
    sd = {k: v.detach().cpu().clone().requires_grad_(v.is_floating_point())
          for k, v in model.state_dict().items()}
    pos_ref = pos0.cpu().clone().requires_grad_(True)
    want, _ = egnn_oracle.model_forward(
        sd, graph.x.cpu(), graph.edge_index.cpu(), pos_ref,
        graph.edge_attr.cpu(), graph.batch.cpu(), num_layers=kw['num_layers'],
        **helpers.oracle_kwargs(kw))
    loss_ref = torch.nn.functional.binary_cross_entropy_with_logits(
        want.reshape(-1), y.cpu())
    loss_ref.backward()
    assert abs(float(loss) - float(loss_ref)) < 1e-5
    for pname, p in model.named_parameters():
        ref = sd[pname].grad
        if ref is None:
            # parameters autograd leaves untouched (e.g. the last layer's
            # coordinate MLP) must stay untouched here too: Adam with weight
            # decay would otherwise move them
            assert p.grad is None, pname
            continue
        assert p.grad is not None, pname
        _close(p.grad.cpu().numpy(), ref.numpy(), pname)
    if pos_ref.grad is not None and kw.get('update_coords', True):
        _close(graph.pos.grad.cpu().numpy(), pos_ref.grad.numpy(), 'pos')
    elif pos_ref.grad is not None:
        _close(graph.pos.grad.cpu().numpy(), pos_ref.grad.numpy(), 'pos')


@pytest.mark.parametrize('math', ['fp32', 'bf16x3'])
def test_backward_is_deterministic(math):
    kw = dict(dim_input=13, dim_output=1, graphnorm=False, **VARIANTS['cfg3_k64'])
    grads = []
    for _ in range(2):
        model = gh.build_model(kw, seed=7, coord_gain=1.0).train()
        model.set_math(math)
        graph = gh.synthetic_graph(900, 2, 400, 15)
        loss = model(graph).sum()
        loss.backward()
        grads.append(torch.cat([p.grad.reshape(-1) for p in model.parameters()
                                if p.grad is not None]))
    assert torch.equal(grads[0], grads[1])


def test_training_step_reduces_loss():
    """A few Adam steps through the reference-style backprop()."""
    kw = dict(dim_input=13, dim_output=1, graphnorm=False, **VARIANTS['cfg3_k64'])
    import pointvs_b200 as pv
    from pathlib import Path
    torch.manual_seed(0)
    model = pv.SartorrasEGNN(Path('/tmp/pvs_test'), 1e-3, 0, None, None,
                             silent=True, **kw).cuda().train()
    graph = gh.synthetic_graph(50, 4, 300, 15)
    graph.y = torch.tensor([1.0, 0.0, 1.0, 0.0], device='cuda')
    graph.lig_fname = graph.rec_fname = ['x'] * 4
    pos = graph.pos.clone()
    losses = []
    for _ in range(6):
        graph.pos = pos.clone()
        y_pred, y_true, _, _ = model.unpack_input_data_and_predict(graph)
        losses.append(model.backprop(y_true, y_pred))
    assert losses[-1] < losses[0]


def test_hub_node_softmax_backward():
    """Softmax attention over a destination with > 128 incoming edges (several
    chunks) differentiates correctly."""
    from oracle import egnn_oracle
    from pointvs_b200 import EGNNLayer
    torch.manual_seed(5)
    n = 300
    layer = EGNNLayer(32, 32, 32, edges_in_d=3, edge_attention=True,
                      normalize=True, tanh=True, softmax_attention=True).cuda()
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
    hc = h.cuda().requires_grad_(True)
    h2, x2, _, m2 = layer(hc, ei.cuda(), x.clone().cuda(), ea.cuda())
    (h2.sum() + x2.sum()).backward()
    sd = {'l.' + k: v.detach().cpu().clone().requires_grad_(True)
          for k, v in layer.state_dict().items()}
    cfg = egnn_oracle.LayerConfig(residual=True, edge_attention=True,
                                  normalize=True, tanh=True,
                                  softmax_attention=True)
    hr = h.clone().requires_grad_(True)
    ho, xo, mo, _ = egnn_oracle.layer_forward(sd, 'l.', cfg, hr, ei[0], ei[1],
                                              x, ea)
    (ho.sum() + xo.sum()).backward()
    _close(hc.grad.cpu().numpy(), hr.grad.numpy(), 'h')
    for pname, p in layer.named_parameters():
        _close(p.grad.cpu().numpy(), sd['l.' + pname].grad.numpy(), pname)


def test_stacked_training_pass_equals_per_layer_pass(monkeypatch):
    """pvs_egnn_stack_fwd / pvs_egnn_stack_bwd (all layers in one call each
    way, PVS_STACK_TRAIN=1) give bit-identical outputs and gradients to the
    per-layer calls."""
    kw = dict(dim_input=13, dim_output=1, graphnorm=False, **VARIANTS['cfg3_k64'])
    results = []
    for stack in ('0', '1'):
        monkeypatch.setenv('PVS_STACK_TRAIN', stack)
        model = gh.build_model(kw, seed=7, coord_gain=1.0).train()
        model.set_math('bf16x3')
        graph = gh.synthetic_graph(900, 2, 400, 15)
        graph.pos = graph.pos.clone().requires_grad_(True)
        out = model(graph)
        out.sum().backward()
        grads = {n: (None if p.grad is None else p.grad.clone())
                 for n, p in model.named_parameters()}
        results.append((out.detach().clone(), graph.pos.grad.clone(), grads))
    (o0, x0, g0), (o1, x1, g1) = results
    assert torch.equal(o0, o1) and torch.equal(x0, x1)
    for name in g0:
        assert (g0[name] is None) == (g1[name] is None), name
        if g0[name] is not None:
            assert torch.equal(g0[name], g1[name]), name


@pytest.mark.parametrize('math', ['fp32', 'bf16x3'])
def test_capacity_bounded_graph_trains_without_a_sync_and_bit_identically(math):
    """edge_capacity='auto': the edge list is allocated at a bound and the true
    edge count never leaves the device (forward, backward and the CSR
    transpose read row_ptr[n]).  Outputs and gradients are bit-identical to
    the exactly-sized graph, and the graph is still un-trimmed afterwards."""
    kw = dict(dim_input=13, dim_output=1, graphnorm=False, **VARIANTS['cfg3_k64'])
    results = []
    for cap in (None, 'auto'):
        model = gh.build_model(kw, seed=7, coord_gain=1.0).train()
        model.set_math(math)
        graph = gh.synthetic_graph(900, 3, 400, 15, ragged=True,
                                   edge_capacity=cap)
        csr = graph.pvs_csr
        assert csr.exact_edge_count == (cap is None)
        graph.pos = graph.pos.clone().requires_grad_(True)
        out = model(graph)
        out.sum().backward()
        if cap is not None:
            assert not csr.exact_edge_count          # nobody forced a read-back
            assert csr.n_edges > csr.true_edge_count()
            csr.check_overflow()
        grads = {n: (None if p.grad is None else p.grad.clone())
                 for n, p in model.named_parameters()}
        results.append((out.detach().clone(), graph.pos.grad.clone(), grads))
    (o0, x0, g0), (o1, x1, g1) = results
    assert torch.equal(o0, o1) and torch.equal(x0, x1)
    for name in g0:
        assert (g0[name] is None) == (g1[name] is None), name
        if g0[name] is not None:
            assert torch.equal(g0[name], g1[name]), name


def test_train_model_raises_on_edge_capacity_overflow():
    """A capacity bound that is too small is reported at the next loss drain
    (not silently trained through)."""
    import pointvs_b200 as pv
    from pathlib import Path
    kw = dict(dim_input=13, dim_output=1, graphnorm=False, **VARIANTS['cfg3_k64'])
    torch.manual_seed(0)
    model = pv.SartorrasEGNN(Path('/tmp/pvs_test_overflow'), 1e-3, 0, None, None,
                             silent=True, **kw).cuda().train()

    def loader(cap):
        g = gh.synthetic_graph(50, 2, 300, 15, edge_capacity=cap)
        g.y = torch.tensor([1.0, 0.0], device='cuda')
        g.lig_fname = g.rec_fname = ['x'] * 2
        return [g]
    losses = model.train_model(loader('auto'), epochs=1)
    assert len(losses) == 1 and np.isfinite(losses[0])
    with pytest.raises(RuntimeError, match='edge capacity exceeded'):
        model.train_model(loader(600), epochs=2)


def test_lean_stacked_pass_equals_per_layer_pass(monkeypatch):
    """The lean training pass (egnn._EGNNStackLeanFn: parameters are not
    autograd inputs, their gradients are accumulated into `p.grad` by the
    backward itself) gives bit-identical outputs and gradients to the
    per-layer pass, also when called twice (gradients accumulate)."""
    kw = dict(dim_input=13, dim_output=1, graphnorm=False, **VARIANTS['cfg3_k64'])
    results = []
    for lean in (False, True):
        monkeypatch.setenv('PVS_STACK_TRAIN', '' if lean else '0')
        model = gh.build_model(kw, seed=7, coord_gain=1.0).train()
        model.set_math('bf16x3')
        model._lean_training = lean
        for _ in range(2):
            graph = gh.synthetic_graph(900, 2, 400, 15, edge_capacity='auto')
            graph.pos = graph.pos.clone().requires_grad_(True)
            out = model(graph)
            out.sum().backward()
        grads = {n: (None if p.grad is None else p.grad.clone())
                 for n, p in model.named_parameters()}
        results.append((out.detach().clone(), graph.pos.grad.clone(), grads))
    (o0, x0, g0), (o1, x1, g1) = results
    assert torch.equal(o0, o1) and torch.equal(x0, x1)
    for name in g0:
        assert (g0[name] is None) == (g1[name] is None), name
        if g0[name] is not None:
            assert torch.equal(g0[name], g1[name]), name


def test_backprop_lean_path_trains_like_the_per_layer_path(monkeypatch):
    """Three `backprop()` steps (arena + fused Adam): the default path (lean
    stacked pass from the second step on) and the per-layer path end with
    bit-identical parameters."""
    import pointvs_b200 as pv
    from pathlib import Path
    kw = dict(dim_input=13, dim_output=1, graphnorm=False, **VARIANTS['cfg3_k64'])
    finals = []
    for env in ('0', ''):
        monkeypatch.setenv('PVS_STACK_TRAIN', env)
        torch.manual_seed(0)
        model = pv.SartorrasEGNN(Path('/tmp/pvs_test_lean'), 1e-3, 1e-4, None, None,
                                 silent=True, **kw).cuda().train()
        model.set_math('bf16x3')
        model.set_record_side_channels(False)
        losses = []
        for s_ in range(3):
            graph = gh.synthetic_graph(50 + s_, 4, 300, 15, edge_capacity='auto')
            graph.y = torch.tensor([1.0, 0.0, 1.0, 0.0], device='cuda')
            graph.lig_fname = graph.rec_fname = ['x'] * 4
            y_pred, y_true, _, _ = model.unpack_input_data_and_predict(graph)
            losses.append(model.backprop(y_true, y_pred))
        assert getattr(model, '_lean_training', False)
        finals.append((losses, [p.detach().clone() for p in model.parameters()]))
    (l0, p0), (l1, p1) = finals
    assert l0 == l1
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)


def test_training_runs_are_bitwise_reproducible():
    """Two identical 12-step `backprop()` runs (ragged batches, capacity-bounded
    graphs, lean stacked pass, gradient reductions on the side stream, launch
    chains, bulk stores) end with bit-identical parameters: no atomics on
    floating-point data and no unordered cross-stream access anywhere."""
    import pointvs_b200 as pv
    from pathlib import Path

    def run():
        kw = dict(dim_input=13, dim_output=1, k=64, num_layers=4,
                  edge_attention=True, node_attention=True, residual=True,
                  normalize=True, tanh=True, graphnorm=False,
                  model_task='classification')
        torch.manual_seed(0)
        m = pv.MultitaskSatorrasEGNN(Path('/tmp/pvs_test_repro'), 1e-3, 1e-4, None,
                                     None, silent=True, **kw).cuda().train()
        m.set_math('bf16x3')
        m.set_record_side_channels(False)
        for s_ in range(12):
            g = gh.synthetic_graph(100 + s_ % 3, 6, 500, 20, ragged=True,
                                   edge_capacity='auto')
            g.y = torch.tensor([float(i % 2) for i in range(6)], device='cuda')
            g.lig_fname = g.rec_fname = [''] * 6
            yp, yt, _, _ = m.unpack_input_data_and_predict(g)
            m.backprop(yt, yp, sync=False)
        torch.cuda.synchronize()
        return torch.cat([p.detach().reshape(-1) for p in m.parameters()])

    a, b = run(), run()
    assert torch.isfinite(a).all()
    assert torch.equal(a, b)
