"""Dense helpers around the EGNN stack: embedding Linear, mean pool, heads.

pvs_linear_fwd replaces `PygLinearPass` (pnn_geometric_base.py:83-94) and the
`feats_linear_layers*` heads (egnn_satorras.py:304-317,
egnn_multitask.py:141-146); pvs_mean_pool_fwd replaces PyG global_mean_pool
(pnn_geometric_base.py:29-33).
"""
import ctypes as C

import torch
from torch import nn

from . import _cabi
from ._cabi import check, lib, ptr, stream


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, weight, bias, act):
        _cabi.require_cuda(inp, weight)
        inp = inp.contiguous().float()
        w = weight.detach().contiguous()
        b = None if bias is None else bias.detach().contiguous()
        rows, ki = inp.shape
        ko = w.shape[0]
        if w.shape[1] != ki:
            raise ValueError(f'linear: input has {ki} features, weight '
                             f'expects {w.shape[1]}')
        out = torch.empty((rows, ko), dtype=torch.float32, device=inp.device)
        with torch.cuda.device(inp.device):
            check(lib().pvs_linear_fwd(
                ptr(inp), ki, rows, ki, ptr(w), ki, ptr(b), ko,
                _cabi.ACT[act], ptr(out), ko, stream()), 'pvs_linear_fwd')
        ctx.act = act
        ctx.has_bias = bias is not None
        ctx.bias_ref = b
        ctx.save_for_backward(inp, w, out)
        return out

    @staticmethod
    def backward(ctx, d_out):
        from .backward import linear_backward
        return linear_backward(ctx, d_out)


def linear(inp, weight, bias, act='none'):
    return _LinearFn.apply(inp, weight, bias, act)


class _MeanPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, graph_ptr, n_graphs):
        h = h.contiguous()
        k = h.shape[1]
        pooled = torch.empty((n_graphs, k), dtype=torch.float32,
                             device=h.device)
        with torch.cuda.device(h.device):
            check(lib().pvs_mean_pool_fwd(ptr(h), ptr(graph_ptr), n_graphs, k,
                                          ptr(pooled), stream()),
                  'pvs_mean_pool_fwd')
        ctx.n_nodes, ctx.k, ctx.n_graphs = h.shape[0], k, n_graphs
        ctx.save_for_backward(graph_ptr)
        return pooled

    @staticmethod
    def backward(ctx, d_pooled):
        from .backward import mean_pool_backward
        return mean_pool_backward(ctx, d_pooled)


def batch_to_ptr(batch, n_graphs):
    batch = batch.to(torch.int64).contiguous()
    out = torch.empty(n_graphs + 1, dtype=torch.int32, device=batch.device)
    with torch.cuda.device(batch.device):
        check(lib().pvs_batch_to_ptr(ptr(batch), batch.numel(), n_graphs,
                                     ptr(out), stream()), 'pvs_batch_to_ptr')
    return out


def mean_pool(h, batch, n_graphs, graph_ptr=None):
    if graph_ptr is None:
        graph_ptr = batch_to_ptr(batch, n_graphs)
    return _MeanPoolFn.apply(h, graph_ptr, n_graphs)


def run_head(seq, feats):
    """Apply an nn.Sequential of Linear / SiLU / ReLU / Softplus modules with
    the activation fused into the preceding Linear."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        if not isinstance(m, nn.Linear):
            raise NotImplementedError(f'head module {type(m).__name__}')
        act = 'none'
        if i + 1 < len(mods) and not isinstance(mods[i + 1], nn.Linear):
            nxt = mods[i + 1]
            act = {nn.SiLU: 'silu', nn.ReLU: 'relu',
                   nn.Softplus: 'softplus'}.get(type(nxt))
            if act is None:
                raise NotImplementedError(f'head module {type(nxt).__name__}')
            i += 1
        feats = linear(feats, m.weight, m.bias, act)
        i += 1
    return feats
