"""Helpers for the -m gpu tests: build the CUDA model from a golden."""
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from tests import helpers


def cuda_model(name, task=None, math='fp32'):
    import pointvs_b200 as pv
    cls, kw, tasks = helpers.MODEL_GOLDENS[name]
    g, sd = helpers.load_model_golden(name)
    C = pv.SartorrasEGNN if cls == 'egnn' else pv.MultitaskSatorrasEGNN
    model = C(Path('/tmp/pvs_test'), 0, 0, None, None, silent=True, **kw)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    model.set_math(math)
    if cls == 'multitask' and task is not None:
        model.set_task(task)
    return model, g


def cuda_graph(g, device='cuda'):
    return SimpleNamespace(
        x=torch.from_numpy(g['in.x']).to(device),
        pos=torch.from_numpy(g['in.pos']).to(device),
        edge_index=torch.from_numpy(g['in.edge_index']).long().to(device),
        edge_attr=torch.from_numpy(g['in.edge_attr']).long().to(device),
        batch=torch.from_numpy(g['in.batch']).long().to(device),
        y=torch.from_numpy(g['in.y']).to(device),
        lig_fname=['lig'] * int(g['in.batch'].max() + 1),
        rec_fname=['rec'] * int(g['in.batch'].max() + 1))


def build_model(kw, multitask=False, seed=0, coord_gain=None):
    import pointvs_b200 as pv
    torch.manual_seed(seed)
    C = pv.MultitaskSatorrasEGNN if multitask else pv.SartorrasEGNN
    model = C(Path('/tmp/pvs_test'), 0, 0, None, None, silent=True, **kw)
    if coord_gain is not None:
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith('coord_mlp.2.weight'):
                    p.mul_(coord_gain / 0.001)
    return model.cuda().eval()


def oracle_forward(model, kw, graph, multitask=False, task='classification',
                   dtype=torch.float32, trace=None):
    """Run the CPU oracle with the CUDA model's weights on the same graph."""
    from oracle import egnn_oracle
    sd = {k: (v.detach().cpu().to(dtype) if v.is_floating_point()
              else v.detach().cpu()) for k, v in model.state_dict().items()}
    out, x = egnn_oracle.model_forward(
        sd, graph.x.cpu().to(dtype), graph.edge_index.cpu(),
        graph.pos.cpu().to(dtype), graph.edge_attr.cpu(), graph.batch.cpu(),
        num_layers=kw['num_layers'], multitask=multitask, model_task=task,
        final_softplus=kw.get('final_softplus', False), trace=trace,
        **helpers.oracle_kwargs(kw))
    return out, x


def synthetic_graph(first_seed, n_complexes, n_atoms, n_lig, radii=(4.0, 4.0),
                    ragged=False, device='cuda', edge_capacity=None):
    """PackedBatch (K1-built CSR) for synthetic complexes."""
    from pointvs_b200.graph import PackedBatch
    from pointvs_b200.synthetic import synthetic_batch
    coords, bp, feats, cptr = synthetic_batch(first_seed, n_complexes, n_atoms,
                                              n_lig, ragged=ragged)
    return PackedBatch.from_arrays(coords, bp, feats, cptr, *radii,
                                   device=device, edge_capacity=edge_capacity)
