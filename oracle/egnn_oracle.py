"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the PointVS EGNN hot path.

A functional, state_dict-driven restatement (plain torch ops on the CPU, fp32
or fp64) of the reference algorithm.  It is the checker for the CUDA path and
the timed "port" for bench.py's cpu_baseline / --impl reference leg.  Only
tests/, __graft_entry__.smoke() and bench.py may import it; nothing under
pointvs_b200/ does.

Parity pin: tests/test_oracle_golden.py checks this file against fixtures in
tests/golden/ that were produced by the *unmodified* reference (imported via
oracle/ref_shim.py by tests/golden/make_golden.py), and
tests/test_oracle_vs_reference.py re-runs that comparison live whenever
/root/reference is present.

Reference lines followed (all under /root/reference/point_vs/models/geometric):
  egnn_satorras.py:178-187  coord2radial      -> _radial
  egnn_satorras.py:123-132  edge_model        -> step "edge MLP"
  egnn_satorras.py:194-202  edge residual     -> step "edge residual"
  egnn_satorras.py:168-176  coord_model       -> step "coordinates"
  egnn_satorras.py:134-166  node_model        -> step "node"
  egnn_satorras.py:332-347  segment sum/mean  -> _seg_sum / _seg_mean
  egnn_satorras.py:319-329  get_embeddings    -> embeddings()
  pnn_geometric_base.py:24-41, 83-94          -> model_forward(), embed
  egnn_multitask.py:96-122, 150-166           -> layer_configs(), heads
Third-party arithmetic restated from published formulas: PyG 2.0.4
global_mean_pool / GraphNorm, torch_scatter scatter_softmax.
"""
from dataclasses import dataclass, replace
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class LayerConfig:
    residual: bool = True
    edge_residual: bool = False
    edge_attention: bool = False
    normalize: bool = False
    tanh: bool = False
    graphnorm: bool = False
    update_coords: bool = True
    permutation_invariance: bool = False
    node_attention: bool = False
    attention_activation_fn: str = 'sigmoid'
    gated_residual: bool = False
    rezero: bool = False
    softmax_attention: bool = False


_ATT_ACT = {
    'sigmoid': torch.sigmoid,
    'tanh': torch.tanh,
    'relu': torch.relu,
    'silu': F.silu,
}


def _seg_sum(data, seg, n):
    out = data.new_zeros((n, data.shape[1]))
    return out.index_add(0, seg, data)


def _seg_mean(data, seg, n):
    total = _seg_sum(data, seg, n)
    count = _seg_sum(torch.ones_like(data), seg, n)
    return total / count.clamp(min=1)


def _seg_softmax(z, seg, n):
    idx = seg.unsqueeze(1)
    mx = z.new_full((n, 1), float('-inf')).scatter_reduce(
        0, idx, z, reduce='amax', include_self=True)
    ex = torch.exp(z - mx[seg])
    den = _seg_sum(ex, seg, n)
    return ex / den[seg]


def _graphnorm(v, w, b, mean_scale, eps=1e-5):
    # Reference calls GraphNorm without `batch` (egnn_satorras.py:84): the
    # statistics run over every node of the mini-batch as if one graph.
    centred = v - v.mean(dim=0, keepdim=True) * mean_scale
    var = (centred * centred).mean(dim=0, keepdim=True)
    return w * centred / torch.sqrt(var + eps) + b


def _radial(x, row, col, normalize):
    diff = x[row] - x[col]
    radial = (diff * diff).sum(dim=1, keepdim=True)
    if normalize:
        diff = diff / (torch.sqrt(radial).detach() + 1e-8)
    return radial, diff


def layer_forward(sd: Dict[str, torch.Tensor], prefix: str, cfg: LayerConfig,
                  h, row, col, x, edge_attr=None, m_prev=None):
    """One EGNNLayer.forward.  Returns (h', x', m, side) with x' a NEW tensor
    (the reference mutates `coord` in place; callers copy back if they need
    that behaviour).  `side` holds att_val / node_att_val tensors."""
    p = lambda name: sd[prefix + name]
    n = h.shape[0]
    side = {}
    radial, diff = _radial(x, row, col, cfg.normalize)

    # edge MLP
    if cfg.permutation_invariance:
        parts = [h[row] + h[col], radial]
    else:
        parts = [h[row], h[col], radial]
    if edge_attr is not None:
        parts.append(edge_attr.to(h.dtype))
    t1 = F.linear(torch.cat(parts, dim=1),
                  p('edge_mlp.0.weight'), p('edge_mlp.0.bias'))
    m = F.silu(F.linear(F.silu(t1),
                        p('edge_mlp.2.weight'), p('edge_mlp.2.bias')))

    # edge residual
    if cfg.edge_residual and m_prev is not None:
        if cfg.rezero:
            m = m_prev + p('edge_gate_parameter') * m
        elif cfg.gated_residual:
            g = torch.relu(p('edge_gate_parameter'))
            m = g * m + (1 - g) * m_prev
        else:
            m = m + m_prev

    # coordinates (uses the pre-attention message)
    if cfg.update_coords:
        q = F.silu(F.linear(m, p('coord_mlp.0.weight'), p('coord_mlp.0.bias')))
        c = F.linear(q, p('coord_mlp.2.weight'))
        if cfg.tanh:
            c = torch.tanh(c)
        x_new = x + _seg_mean(diff * c, row, n)
    else:
        x_new = x

    # node
    if cfg.edge_attention:
        z = F.linear(m, p('att_mlp.0.weight'), p('att_mlp.0.bias'))
        if cfg.softmax_attention:
            att = _seg_softmax(z, row, n)
        else:
            att = _ATT_ACT[cfg.attention_activation_fn](z)
        side['att_val'] = att
        agg = _seg_sum(att * m, row, n)
    else:
        agg = _seg_sum(m, row, n)
    v = F.linear(torch.cat([h, agg], dim=1),
                 p('node_mlp.0.weight'), p('node_mlp.0.bias'))
    if cfg.graphnorm:
        v = _graphnorm(v, p('node_mlp.1.weight'), p('node_mlp.1.bias'),
                       p('node_mlp.1.mean_scale'))
    o = F.linear(F.silu(v), p('node_mlp.3.weight'), p('node_mlp.3.bias'))
    if cfg.node_attention:
        natt = F.linear(o, p('node_att_mlp.0.weight'),
                        p('node_att_mlp.0.bias'))
        natt = natt if cfg.softmax_attention else \
            _ATT_ACT[cfg.attention_activation_fn](natt)
        o = o * natt
        side['node_att_val'] = natt
    if cfg.residual:
        if cfg.rezero:
            o = h + p('node_gate_parameter') * o
        elif cfg.gated_residual:
            g = torch.relu(p('node_gate_parameter'))
            o = g * o + (1 - g) * h
        else:
            o = h + o
    return o, x_new, m, side


def layer_configs(num_layers: int, multitask: bool = False,
                  node_attention_final_only=False,
                  edge_attention_final_only=False,
                  node_attention_first_only=False,
                  edge_attention_first_only=False,
                  **kw) -> List[LayerConfig]:
    """Per-layer flags.  build_net defaults of the MODEL classes apply
    (normalize=True, tanh=True, graphnorm=True: egnn_satorras.py:212-238),
    and the multitask class thins attention per layer
    (egnn_multitask.py:96-122)."""
    base = LayerConfig(
        residual=kw.get('residual', True),
        edge_residual=kw.get('edge_residual', False),
        edge_attention=kw.get('edge_attention', False),
        normalize=kw.get('normalize', True),
        tanh=kw.get('tanh', True),
        graphnorm=kw.get('graphnorm', True),
        update_coords=kw.get('update_coords', True),
        permutation_invariance=kw.get('permutation_invariance', False),
        node_attention=kw.get('node_attention', False),
        attention_activation_fn=kw.get('attention_activation_fn', 'sigmoid'),
        gated_residual=kw.get('gated_residual', False),
        rezero=kw.get('rezero', False),
        softmax_attention=kw.get('softmax_attention', False))
    out = []
    for i in range(num_layers):
        cfg = base
        if multitask:
            def keep(on, first_only, final_only):
                if not on:
                    return False
                if not first_only and not final_only:
                    return True
                if first_only and i == 0:
                    return True
                return bool(final_only and i == num_layers - 1)
            cfg = replace(
                cfg,
                node_attention=keep(base.node_attention,
                                    node_attention_first_only,
                                    node_attention_final_only),
                edge_attention=keep(base.edge_attention,
                                    edge_attention_first_only,
                                    edge_attention_final_only))
        out.append(cfg)
    return out


def embeddings(sd, cfgs: List[LayerConfig], feats, edge_index, coords,
               edge_attr, trace: Optional[list] = None):
    """get_embeddings: layers.0 is the Linear embed, layers.1..L are EGNN."""
    row, col = edge_index[0], edge_index[1]
    h = F.linear(feats, sd['layers.0.m.weight'], sd['layers.0.m.bias'])
    x, m = coords, None
    for i, cfg in enumerate(cfgs):
        h, x, m, side = layer_forward(
            sd, f'layers.{i + 1}.', cfg, h, row, col, x, edge_attr, m)
        if trace is not None:
            trace.append({'h': h, 'x': x, 'm': m, **side})
    return h, x, m


def mean_pool(h, batch, size):
    if size == 1:
        return h.mean(dim=0)
    tot = h.new_zeros((size, h.shape[1])).index_add(0, batch, h)
    cnt = h.new_zeros((size,)).index_add(
        0, batch, torch.ones_like(batch, dtype=h.dtype))
    return tot / cnt.clamp(min=1).unsqueeze(1)


def _head(sd, prefix, feats, final_act=None):
    idx = sorted({int(k[len(prefix):].split('.')[0]) for k in sd
                  if k.startswith(prefix) and k.endswith('.weight')})
    for n, i in enumerate(idx):
        feats = F.linear(feats, sd[f'{prefix}{i}.weight'],
                         sd[f'{prefix}{i}.bias'])
        if n < len(idx) - 1:
            feats = F.silu(feats)
    if final_act == 'softplus':
        feats = F.softplus(feats)
    elif final_act == 'relu':
        feats = torch.relu(feats)
    return feats


def model_forward(sd, feats, edge_index, coords, edge_attr, batch,
                  num_layers, multitask=False, model_task='classification',
                  final_softplus=False, trace=None, **model_kwargs):
    """SartorrasEGNN.forward / MultitaskSatorrasEGNN.forward on unpacked
    tensors.  Returns (logits, final coords)."""
    cfgs = layer_configs(num_layers, multitask=multitask, **model_kwargs)
    size = int(batch.max()) + 1
    h, x, _ = embeddings(sd, cfgs, feats, edge_index, coords, edge_attr,
                         trace=trace)
    pooled = mean_pool(h, batch, size)
    if multitask:
        if 'classification' in model_task:
            out = _head(sd, 'feats_linear_layers_pose.', pooled)
        else:
            out = _head(sd, 'feats_linear_layers_affinity.', pooled,
                        'softplus' if final_softplus else 'relu')
    else:
        out = _head(sd, 'feats_linear_layers.', pooled,
                    'softplus' if final_softplus else None)
    return out, x
