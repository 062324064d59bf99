"""TEST INFRASTRUCTURE ONLY -- the reference's own CPU path, timed by bench.py.

`bench.py --impl reference` and the `cpu_baseline` leg call this module; it runs
the UNMODIFIED reference classes (`generate_edges`, the Dataset's graph
assembly, PyG-style collate, `SartorrasEGNN.forward`) imported through
oracle/ref_shim.py from /root/reference (build container) or from the
untracked copy oracle/_ref (GPU box; oracle/make_ref.py).  When neither
exists `available()` is False and bench.py falls back to the oracle port.

Mirrors SURVEY.md 8(d) "CPU baseline beside it": all host threads for the
model, `model.eval()`, `torch.no_grad()`, graph building in
min(4, cores) worker processes like the reference's DataLoader
(global_objects.py:25 NUM_WORKERS).  Nothing under pointvs_b200/ imports this.
"""
import os
import time

import numpy as np

from oracle import ref_shim

_POOL = None
_REF = None


def available():
    return ref_shim.reference_available()


def _ref():
    global _REF
    if _REF is None:
        _REF = ref_shim.import_reference()
    return _REF


def _worker_init():
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    _ref()


def _graph_of(args):
    """One complex -> reference `Data` fields, as PygPointCloudDataset.__getitem__
    assembles them (data_loaders.py:359-391 of the reference)."""
    import pandas as pd
    coords, bp, feats, inter, intra = args
    struct = pd.DataFrame({'x': coords[:, 0], 'y': coords[:, 1],
                           'z': coords[:, 2], 'bp': bp})
    _, (row, col), attr = _ref().generate_edges(
        struct, inter_radius=inter, intra_radius=intra, prune=False)
    return (np.vstack([np.asarray(row), np.asarray(col)]).astype(np.int64),
            np.asarray(attr).astype(np.int64))


def workers(cores=None):
    cores = cores or os.cpu_count() or 1
    return max(1, min(4, cores))      # global_objects.py:25 of the reference


def _pool():
    global _POOL
    if _POOL is None:
        import multiprocessing as mp
        _POOL = mp.get_context('spawn').Pool(workers(), initializer=_worker_init)
    return _POOL


def close():
    global _POOL
    if _POOL is not None:
        _POOL.terminate()
        _POOL = None


def build_model(model_kw, state_dict=None, seed=0):
    """The reference's SartorrasEGNN on the CPU (its own default init under
    torch.manual_seed(seed), or `state_dict` -- the key names are the
    reference's, so a state_dict of the CUDA model loads unchanged)."""
    import tempfile
    from pathlib import Path
    import torch
    torch.manual_seed(seed)
    tmp = Path(tempfile.mkdtemp(prefix='pvs_ref_'))
    model = _ref().SartorrasEGNN(tmp, 0, 0, None, None, silent=True, **model_kw)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    return model.eval()


def step(model, complexes, inter, intra, parallel=True):
    """Graph build + collate + forward for a list of (coords, bp, feats).
    Returns (scores [B], n_edges, t_graph_s, t_model_s)."""
    import torch
    from torch.nn.functional import one_hot
    ref = _ref()
    t0 = time.perf_counter()
    jobs = [(c, b, f, inter, intra) for c, b, f in complexes]
    if parallel and workers() > 1 and len(jobs) > 1:
        graphs = _pool().map(_graph_of, jobs, chunksize=max(1, len(jobs) // (4 * workers())))
    else:
        graphs = [_graph_of(j) for j in jobs]
    items = []
    for (coords, bp, feats), (ei, attr) in zip(complexes, graphs):
        items.append(ref.Data(
            x=torch.from_numpy(feats), edge_index=torch.from_numpy(ei).long(),
            edge_attr=one_hot(torch.from_numpy(attr).long(), 3),
            pos=torch.from_numpy(coords.astype(np.float32)),
            y=torch.tensor(0).long(), rec_fname='rec', lig_fname='lig'))
    graph = ref.collate(items)
    t_graph = time.perf_counter() - t0
    t0 = time.perf_counter()
    with torch.no_grad():
        # forward mutates graph.pos in place (egnn_satorras.py:174): the graph
        # is rebuilt every step, so no clone is needed here
        out = model(graph)
    t_model = time.perf_counter() - t0
    return out.reshape(-1), int(graph.edge_index.shape[1]), t_graph, t_model
