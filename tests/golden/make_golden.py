"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every array written here is produced by the reference's own code
(`generate_edges`, `SartorrasEGNN`, `MultitaskSatorrasEGNN`, its Dataset)
imported through oracle/ref_shim.py; nothing from this repository's CUDA path
or oracle takes part.  The fixtures pin the oracle (tests/test_oracle_golden.py)
and are compared directly with the CUDA path (tests/test_gpu_*.py).
"""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import pandas as pd
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))

from oracle import ref_shim  # noqa: E402
from pointvs_b200.synthetic import synthetic_complex  # noqa: E402

REF = ref_shim.import_reference()


# --------------------------------------------------------------------------
# edges
# --------------------------------------------------------------------------
def ref_edges(coords, bp, inter, intra, prune=False):
    struct = pd.DataFrame({'x': coords[:, 0], 'y': coords[:, 1],
                           'z': coords[:, 2], 'bp': bp})
    out_struct, (row, col), attr = REF.generate_edges(
        struct, inter_radius=inter, intra_radius=intra, prune=prune)
    kept = out_struct.index.to_numpy() if not prune else None
    return np.asarray(row), np.asarray(col), np.asarray(attr), len(out_struct)


def make_edge_goldens():
    out = {}
    cases = {
        'syn300_r4_r4': (synthetic_complex(7, 300, 20), 4.0, 4.0),
        'syn300_r4_r2': (synthetic_complex(8, 300, 20), 4.0, 2.0),
        'syn257_r6_r2': (synthetic_complex(9, 257, 11), 6.0, 2.0),
        'syn64_r10_r3': (synthetic_complex(10, 64, 5), 10.0, 3.0),
    }
    for name, ((coords, bp, _), inter, intra) in cases.items():
        row, col, attr, _ = ref_edges(coords, bp, inter, intra)
        out[name + '.coords'] = coords
        out[name + '.bp'] = bp
        out[name + '.radii'] = np.array([inter, intra])
        out[name + '.row'] = row.astype(np.int32)
        out[name + '.col'] = col.astype(np.int32)
        out[name + '.attr'] = attr.astype(np.int8)
    # coincident atoms + an isolated far atom + boundary-exact distances
    coords = np.array([[0, 0, 0], [0, 0, 0], [4, 0, 0], [0, 3.999999, 0],
                       [100, 100, 100], [2, 0, 0], [0, 0, 4.000001],
                       [1e-8, 0, 0]], dtype=np.float64)
    bp = np.array([0, 1, 1, 0, 1, 0, 1, 1], dtype=np.int32)
    row, col, attr, _ = ref_edges(coords, bp, 4.0, 2.0)
    out['edgecases.coords'], out['edgecases.bp'] = coords, bp
    out['edgecases.radii'] = np.array([4.0, 2.0])
    out['edgecases.row'] = row.astype(np.int32)
    out['edgecases.col'] = col.astype(np.int32)
    out['edgecases.attr'] = attr.astype(np.int8)
    # prune: two blobs far apart; only the one holding the ligand survives
    c1, b1, _ = synthetic_complex(11, 120, 10)
    c2, _, _ = synthetic_complex(12, 40, 0)
    coords = np.concatenate([c1, c2 + 60.0])
    bp = np.concatenate([b1, np.ones(40, dtype=np.int32)])
    row, col, attr, n_kept = ref_edges(coords, bp, 4.0, 2.0, prune=True)
    out['prune.coords'], out['prune.bp'] = coords, bp
    out['prune.radii'] = np.array([4.0, 2.0])
    out['prune.row'] = row.astype(np.int32)
    out['prune.col'] = col.astype(np.int32)
    out['prune.attr'] = attr.astype(np.int8)
    out['prune.n_kept'] = np.array([n_kept])
    np.savez_compressed(HERE / 'edges.npz', **out)
    print('edges.npz:', {k: v.shape for k, v in out.items()
                         if k.endswith('.row')})


# --------------------------------------------------------------------------
# the reference's own 82-node test graph (test/setup_and_params.py:15-27)
# --------------------------------------------------------------------------
def make_fixture82():
    cwd = os.getcwd()
    os.chdir(ref_shim.REFERENCE_ROOT)
    try:
        from point_vs.preprocessing.data_loaders import (
            get_data_loader, PygPointCloudDataset)
        dl = get_data_loader(
            Path('test/resources'), dataset_class=PygPointCloudDataset,
            batch_size=2, compact=True, radius=4, use_atomic_numbers=False,
            rot=False, augmented_actives=0, min_aug_angle=0,
            polar_hydrogens=False, receptors=None, mode='val',
            types_fname=Path('test/resources/test.types'),
            fname_suffix='.parquet', edge_radius=4, estimate_bonds=True)
        dl.num_workers = 0
        g = next(iter(torch.utils.data.DataLoader(
            dl.dataset, batch_size=2, shuffle=False,
            collate_fn=REF.collate)))
    finally:
        os.chdir(cwd)
    out = {
        'x': g.x.numpy().astype(np.float32),
        'pos': g.pos.numpy().astype(np.float32),
        'edge_index': g.edge_index.numpy().astype(np.int32),
        'edge_attr': g.edge_attr.numpy().astype(np.int8),
        'batch': g.batch.numpy().astype(np.int32),
    }
    np.savez_compressed(HERE / 'fixture82.npz', **out)
    print('fixture82.npz:', {k: v.shape for k, v in out.items()})
    return g


# --------------------------------------------------------------------------
# model goldens
# --------------------------------------------------------------------------
def synthetic_graph(seeds, n_atoms, n_lig, inter, intra):
    items = []
    for s in seeds:
        coords, bp, feats = synthetic_complex(s, n_atoms, n_lig)
        row, col, attr, _ = ref_edges(coords, bp, inter, intra)
        items.append(REF.Data(
            x=torch.from_numpy(feats),
            edge_index=torch.from_numpy(np.vstack([row, col])).long(),
            edge_attr=torch.nn.functional.one_hot(
                torch.from_numpy(attr).long(), 3),
            pos=torch.from_numpy(coords).float(),
            y=torch.tensor(float(s % 2))))
    return REF.collate(items)


def run_reference_model(model, graph, multitask):
    """Layer-by-layer replica of get_embeddings so per-layer state is kept.
    Uses the reference's own modules; clones pos (forward mutates it)."""
    feats, edges, coords, eattr, batch = model.unpack_graph(graph)
    coords = coords.clone()
    trace = []
    messages = None
    for layer in model.layers:
        feats, coords, eattr, messages = layer(
            h=feats, edge_index=edges, coord=coords, edge_attr=eattr,
            edge_messages=messages)
        rec = {'h': feats.detach().clone(), 'x': coords.detach().clone()}
        if messages is not None:
            rec['m'] = messages.detach().clone()
        if getattr(layer, 'att_val', None) is not None:
            rec['att'] = torch.from_numpy(np.asarray(layer.att_val))
        if getattr(layer, 'node_att_val', None) is not None:
            rec['natt'] = torch.from_numpy(np.asarray(layer.node_att_val))
        trace.append(rec)
    return trace


def make_model_golden(name, graph, cls_name, model_kwargs, seed,
                      coord_gain=1.0, randomise_gates=False, with_grads=False,
                      tasks=('classification',)):
    torch.manual_seed(seed)
    cls = getattr(REF, cls_name)
    multitask = cls_name == 'MultitaskSatorrasEGNN'
    with tempfile.TemporaryDirectory() as tmp:
        model = cls(Path(tmp), 0, 0, None, None, silent=True, **model_kwargs)
    model.eval()
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for pname, p in model.named_parameters():
            if pname.endswith('coord_mlp.2.weight'):
                p.mul_(coord_gain / 0.001)
            if randomise_gates and 'gate_parameter' in pname:
                p.copy_(torch.rand(1, generator=gen) * 0.8 + 0.1)
            if pname.endswith('node_mlp.1.weight') or \
                    pname.endswith('node_mlp.1.mean_scale'):
                p.copy_(torch.rand(p.shape, generator=gen) + 0.5)
            if pname.endswith('node_mlp.1.bias'):
                p.copy_(torch.rand(p.shape, generator=gen) - 0.5)
    out = {'sd.' + k: v.detach().numpy().copy()
           for k, v in model.state_dict().items()}
    out['in.x'] = graph.x.numpy().astype(np.float32)
    out['in.pos'] = graph.pos.numpy().astype(np.float32)
    out['in.edge_index'] = graph.edge_index.numpy().astype(np.int32)
    out['in.edge_attr'] = graph.edge_attr.numpy().astype(np.int8)
    out['in.batch'] = graph.batch.numpy().astype(np.int32)
    out['in.y'] = torch.as_tensor(graph.y).float().numpy()

    def fresh():
        g = REF.Data(**{k: (getattr(graph, k).clone()
                            if isinstance(getattr(graph, k), torch.Tensor)
                            else getattr(graph, k)) for k in graph.keys()})
        return g

    with torch.no_grad():
        trace = run_reference_model(model, fresh(), multitask)
    for i, rec in enumerate(trace):
        for k, v in rec.items():
            if k == 'm' and i != len(trace) - 1:
                continue   # [E,k] per layer is bulky: keep the last only
            out[f'layer{i}.{k}'] = v.numpy().astype(np.float32)
    for task in tasks:
        if multitask:
            model.set_task(task)
        with torch.no_grad():
            logits = model(fresh())
        out[f'out.{task}'] = logits.detach().numpy().astype(np.float32)
    if with_grads:
        if multitask:
            model.set_task('classification')
        model.zero_grad()
        g = fresh()
        pos_leaf = g.pos.clone().requires_grad_(False)
        logits = model(g).reshape(-1)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(
            logits, torch.as_tensor(graph.y).float().reshape(-1))
        loss.backward()
        out['grad.loss'] = np.array([float(loss)], dtype=np.float32)
        for pname, p in model.named_parameters():
            if p.grad is not None:
                out['grad.' + pname] = p.grad.numpy().astype(np.float32)
    np.savez_compressed(HERE / f'model_{name}.npz', **out)
    size = os.path.getsize(HERE / f'model_{name}.npz')
    print(f'model_{name}.npz: {size / 1e3:.0f} kB, logits',
          {t: out[f"out.{t}"].reshape(-1)[:4] for t in tasks})


def main():
    make_edge_goldens()
    g82 = make_fixture82()

    g_small = synthetic_graph([21, 22, 23], 90, 9, 4.0, 4.0)
    g_bonds = synthetic_graph([31, 32], 120, 12, 4.0, 2.0)   # degree-0 nodes
    g_k64 = synthetic_graph([41, 42], 110, 10, 4.0, 4.0)

    cfg3 = dict(dim_input=13, dim_output=1, edge_attention=True,
                node_attention=True, residual=True, normalize=True, tanh=True,
                graphnorm=False)
    make_model_golden('cfg3_k32', g_small, 'SartorrasEGNN',
                      dict(cfg3, k=32, num_layers=3), seed=0,
                      with_grads=True)
    make_model_golden('cfg3_k64_l8', g_k64, 'SartorrasEGNN',
                      dict(cfg3, k=64, num_layers=8), seed=0)
    make_model_golden('cfg3_k32_gain001', g_small, 'SartorrasEGNN',
                      dict(cfg3, k=32, num_layers=2), seed=3,
                      coord_gain=0.001)
    make_model_golden('alloff_multitask', g_bonds, 'MultitaskSatorrasEGNN',
                      dict(dim_input=13, dim_output=1, k=32, num_layers=3,
                           edge_attention=False, node_attention=False,
                           residual=False, normalize=False, tanh=False,
                           graphnorm=False, model_task='classification'),
                      seed=1, with_grads=True,
                      tasks=('classification', 'regression'))
    g82.y = torch.tensor([1.0, 0.0])
    make_model_golden('testkwargs_fixture82', g82, 'SartorrasEGNN',
                      dict(k=32, num_layers=3, dropout=0, dim_input=12,
                           dim_output=1, graphnorm=True, update_coords=True,
                           node_attention=True, residual=True,
                           edge_attention=True, softmax_attention=True,
                           cache=False, dim_hidden=32, pooling_only=True),
                      seed=2, with_grads=True)
    make_model_golden('gated_tanhatt_multifc', g_bonds, 'SartorrasEGNN',
                      dict(dim_input=13, dim_output=3, k=32, num_layers=3,
                           edge_attention=True, node_attention=True,
                           residual=True, edge_residual=True,
                           gated_residual=True, normalize=True, tanh=False,
                           graphnorm=False, attention_activation_fn='tanh',
                           multi_fc=True, final_softplus=True,
                           model_task='multi_regression'),
                      seed=4, randomise_gates=True)
    make_model_golden('rezero_perminv_static', g_small, 'SartorrasEGNN',
                      dict(dim_input=13, dim_output=1, k=16, num_layers=2,
                           edge_attention=True, node_attention=False,
                           residual=True, edge_residual=True, rezero=True,
                           normalize=False, tanh=True, graphnorm=False,
                           attention_activation_fn='silu',
                           permutation_invariance=True, update_coords=False),
                      seed=5, randomise_gates=True)
    make_model_golden('multitask_firstfinal', g_small,
                      'MultitaskSatorrasEGNN',
                      dict(dim_input=13, dim_output=1, k=48, num_layers=3,
                           edge_attention=True, node_attention=True,
                           edge_attention_first_only=True,
                           node_attention_final_only=True,
                           residual=True, edge_residual=True,
                           normalize=True, tanh=True,
                           graphnorm=False, final_softplus=True,
                           attention_activation_fn='relu',
                           model_task='regression'),
                      seed=6, tasks=('regression', 'classification'))


if __name__ == '__main__':
    main()
