"""Shared test helpers: golden loading and oracle drivers."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

MODEL_GOLDENS = {
    # name: (class, model_kwargs, tasks)
    'cfg3_k32': ('egnn', dict(
        dim_input=13, dim_output=1, k=32, num_layers=3, edge_attention=True,
        node_attention=True, residual=True, normalize=True, tanh=True,
        graphnorm=False), ('classification',)),
    'cfg3_k64_l8': ('egnn', dict(
        dim_input=13, dim_output=1, k=64, num_layers=8, edge_attention=True,
        node_attention=True, residual=True, normalize=True, tanh=True,
        graphnorm=False), ('classification',)),
    'cfg3_k32_gain001': ('egnn', dict(
        dim_input=13, dim_output=1, k=32, num_layers=2, edge_attention=True,
        node_attention=True, residual=True, normalize=True, tanh=True,
        graphnorm=False), ('classification',)),
    'alloff_multitask': ('multitask', dict(
        dim_input=13, dim_output=1, k=32, num_layers=3, edge_attention=False,
        node_attention=False, residual=False, normalize=False, tanh=False,
        graphnorm=False, model_task='classification'),
        ('classification', 'regression')),
    'testkwargs_fixture82': ('egnn', dict(
        k=32, num_layers=3, dropout=0, dim_input=12, dim_output=1,
        graphnorm=True, update_coords=True, node_attention=True,
        residual=True, edge_attention=True, softmax_attention=True,
        cache=False, dim_hidden=32, pooling_only=True), ('classification',)),
    'gated_tanhatt_multifc': ('egnn', dict(
        dim_input=13, dim_output=3, k=32, num_layers=3, edge_attention=True,
        node_attention=True, residual=True, edge_residual=True,
        gated_residual=True, normalize=True, tanh=False, graphnorm=False,
        attention_activation_fn='tanh', multi_fc=True, final_softplus=True,
        model_task='multi_regression'), ('classification',)),
    'rezero_perminv_static': ('egnn', dict(
        dim_input=13, dim_output=1, k=16, num_layers=2, edge_attention=True,
        node_attention=False, residual=True, edge_residual=True, rezero=True,
        normalize=False, tanh=True, graphnorm=False,
        attention_activation_fn='silu', permutation_invariance=True,
        update_coords=False), ('classification',)),
    'multitask_firstfinal': ('multitask', dict(
        dim_input=13, dim_output=1, k=48, num_layers=3, edge_attention=True,
        node_attention=True, edge_attention_first_only=True,
        node_attention_final_only=True, residual=True, edge_residual=True,
        normalize=True, tanh=True, graphnorm=False, final_softplus=True,
        attention_activation_fn='relu', model_task='regression'),
        ('regression', 'classification')),
}


def load_npz(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def load_model_golden(name):
    g = load_npz(f'model_{name}.npz')
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items()
          if k.startswith('sd.')}
    return g, sd


def oracle_kwargs(model_kwargs):
    skip = {'dim_input', 'dim_output', 'k', 'num_layers', 'dropout', 'cache',
            'dim_hidden', 'pooling_only', 'multi_fc', 'model_task',
            'final_softplus'}
    return {k: v for k, v in model_kwargs.items() if k not in skip}


def run_oracle(name, task=None, dtype=torch.float32, trace=None):
    from oracle import egnn_oracle
    cls, kw, tasks = MODEL_GOLDENS[name]
    g, sd = load_model_golden(name)
    sd = {k: v.to(dtype) if v.is_floating_point() else v
          for k, v in sd.items()}
    task = task or tasks[0]
    ei = torch.from_numpy(g['in.edge_index']).long()
    out, x = egnn_oracle.model_forward(
        sd, torch.from_numpy(g['in.x']).to(dtype), ei,
        torch.from_numpy(g['in.pos']).to(dtype),
        torch.from_numpy(g['in.edge_attr']).long(),
        torch.from_numpy(g['in.batch']).long(),
        num_layers=kw['num_layers'], multitask=(cls == 'multitask'),
        model_task=task if cls == 'multitask' else kw.get(
            'model_task', 'classification'),
        final_softplus=kw.get('final_softplus', False), trace=trace,
        **oracle_kwargs(kw))
    return g, out, x


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30)))


def scaled_err(a, b):
    """max |a-b| / max|b|: for per-element arrays that cross zero."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))
