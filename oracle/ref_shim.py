"""TEST INFRASTRUCTURE ONLY -- import shim for the real PointVS reference.

This module makes the *unmodified* reference under ``/root/reference``
importable in the build container, where eight of its third-party
dependencies (torch_scatter, torch_geometric, matplotlib, pymol, rdkit,
egnn_pytorch, plip, openbabel) are absent.  It is used for exactly two things:

  * ``tests/golden/make_golden.py`` -- generating the committed golden
    fixtures that pin ``oracle/egnn_oracle.py`` and ``oracle/radius_graph.py``
    to the reference's own outputs;
  * ``tests/test_oracle_vs_reference.py`` -- a live cross-check that is
    skipped when ``/root/reference`` does not exist (i.e. on the GPU box);
  * ``bench.py``'s CPU legs (``--impl reference`` and ``cpu_baseline``), which
    time the reference's own classes on the host cores; on the GPU box they
    import the untracked copy under ``oracle/_ref`` (``oracle/make_ref.py``).

Nothing in ``pointvs_b200/`` may import this file.  The four third-party
functions the EGNN path actually executes are restated from their published
formulas (PyG 2.0.4 ``global_mean_pool`` / ``GraphNorm`` / collate,
torch_scatter ``scatter_softmax``); every other attribute of the missing
packages stays a ``MagicMock``.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest import mock

# The reference's own tree when present (build container); otherwise the
# untracked copy of its package made by oracle/make_ref.py (travels to the GPU
# box with the gpurun snapshot: oracle/_ref is git-ignored, not gpurun-ignored).
_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = '/root/reference'
if not os.path.isdir(os.path.join(REFERENCE_ROOT, 'point_vs')):
    REFERENCE_ROOT = os.path.join(_HERE, '_ref')
_MISSING_ROOTS = ('torch_scatter', 'torch_geometric', 'matplotlib', 'pymol',
                  'rdkit', 'egnn_pytorch', 'plip', 'openbabel', 'wandb')


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'point_vs'))


class _MockLoader(importlib.abc.Loader):
    def create_module(self, spec):
        mod = mock.MagicMock(name=spec.name)
        mod.__name__ = spec.name
        mod.__path__ = []
        mod.__spec__ = spec
        mod.__loader__ = self
        return mod

    def exec_module(self, module):
        return None


class _MockFinder(importlib.abc.MetaPathFinder):
    def __init__(self, roots):
        self.roots = set(roots)
        self.loader = _MockLoader()

    def find_spec(self, name, path=None, target=None):
        if name.split('.')[0] in self.roots:
            return importlib.machinery.ModuleSpec(
                name, self.loader, is_package=True)
        return None


def _scatter_softmax(src, index, dim=0):
    """torch_scatter.composite.scatter_softmax: max-shifted softmax per group."""
    import torch
    n = int(index.max()) + 1 if index.numel() else 0
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    mx = torch.full((n,) + tuple(src.shape[1:]), float('-inf'),
                    dtype=src.dtype, device=src.device)
    mx = mx.scatter_reduce(0, idx, src, reduce='amax', include_self=True)
    ex = torch.exp(src - mx.gather(0, idx))
    den = torch.zeros_like(mx).scatter_add_(0, idx, ex)
    return ex / den.gather(0, idx)


def _global_mean_pool(x, batch, size=None):
    import torch
    size = int(batch.max()) + 1 if size is None else int(size)
    out = torch.zeros(size, x.shape[1], dtype=x.dtype, device=x.device)
    out.index_add_(0, batch, x)
    cnt = torch.zeros(size, dtype=x.dtype, device=x.device)
    cnt.index_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return out / cnt.clamp(min=1).unsqueeze(1)


def _make_graphnorm():
    import torch
    from torch import nn

    class GraphNorm(nn.Module):
        """PyG 2.0.4 GraphNorm formula."""

        def __init__(self, in_channels, eps=1e-5):
            super().__init__()
            self.in_channels = in_channels
            self.eps = eps
            self.weight = nn.Parameter(torch.ones(in_channels))
            self.bias = nn.Parameter(torch.zeros(in_channels))
            self.mean_scale = nn.Parameter(torch.ones(in_channels))

        def forward(self, x, batch=None):
            if batch is None:
                batch = x.new_zeros(x.size(0), dtype=torch.long)
            size = int(batch.max()) + 1
            mean = _global_mean_pool(x, batch, size)
            out = x - mean.index_select(0, batch) * self.mean_scale
            var = _global_mean_pool(out * out, batch, size)
            std = (var + self.eps).sqrt().index_select(0, batch)
            return self.weight * out / std + self.bias

    return GraphNorm


class _Data:
    """Attribute bag standing in for torch_geometric.data.Data."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith('_')]

    def to(self, device):
        import torch
        for k in self.keys():
            v = getattr(self, k)
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device))
        return self


def _collate(items):
    """PyG Batch.from_data_list semantics for the fields the reference uses."""
    import torch
    out = _Data()
    keys = items[0].keys()
    offsets, n = [], 0
    for it in items:
        offsets.append(n)
        n += it.x.shape[0] if hasattr(it, 'x') else it.pos.shape[0]
    for k in keys:
        vals = [getattr(it, k) for it in items]
        if isinstance(vals[0], torch.Tensor):
            if k == 'edge_index':
                setattr(out, k, torch.cat(
                    [v + o for v, o in zip(vals, offsets)], dim=1))
            elif vals[0].dim() == 0:
                setattr(out, k, torch.stack(vals))
            else:
                setattr(out, k, torch.cat(vals, dim=0))
        else:
            setattr(out, k, list(vals))
    sizes = [it.x.shape[0] for it in items]
    out.batch = torch.repeat_interleave(
        torch.arange(len(items)), torch.tensor(sizes))
    return out


_INSTALLED = False


def install():
    """Install the shim (idempotent) and put the reference on sys.path."""
    global _INSTALLED
    if _INSTALLED:
        return
    if not reference_available():
        raise RuntimeError('reference tree not present at ' + REFERENCE_ROOT)
    import numpy
    import torch
    from torch import nn
    import torch.utils.data as tud

    os.environ.setdefault('WANDB_MODE', 'disabled')
    sys.meta_path.insert(0, _MockFinder(_MISSING_ROOTS))
    sys.path.insert(0, REFERENCE_ROOT)
    if not hasattr(numpy, 'product'):
        numpy.product = numpy.prod

    import torch_scatter
    torch_scatter.composite.scatter_softmax = _scatter_softmax

    import torch_geometric
    import torch_geometric.nn
    import torch_geometric.nn.norm
    import torch_geometric.data
    import torch_geometric.loader
    import torch_geometric.utils
    graphnorm = _make_graphnorm()
    torch_geometric.nn.global_mean_pool = _global_mean_pool
    torch_geometric.nn.GraphNorm = graphnorm
    torch_geometric.nn.norm.GraphNorm = graphnorm

    class MessagePassing(nn.Module):
        def __init__(self, *args, **kwargs):
            super().__init__()

    torch_geometric.nn.MessagePassing = MessagePassing
    torch_geometric.data.Data = _Data
    torch_geometric.data.Dataset = tud.Dataset

    class DataLoader(tud.DataLoader):
        def __init__(self, dataset, batch_size=1, shuffle=False, **kwargs):
            kwargs.pop('collate_fn', None)
            super().__init__(dataset, batch_size=batch_size, shuffle=shuffle,
                             collate_fn=_collate, **kwargs)

    torch_geometric.loader.DataLoader = DataLoader

    import wandb
    wandb.watch = lambda *a, **k: None
    _INSTALLED = True


def import_reference():
    """Return a namespace holding the reference symbols on the hot path.

    The reference chooses its DEVICE once, at import (global_objects.py:14-22:
    CUDA when available).  The oracle and bench.py's CPU legs want its CPU path
    -- also on the GPU box -- so CUDA is reported unavailable while its modules
    are imported; nothing else about the reference is changed."""
    install()
    import warnings
    import torch
    real_avail = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    try:
        return _import_reference(warnings)
    finally:
        torch.cuda.is_available = real_avail


def _import_reference(warnings):
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        from point_vs.models.geometric.egnn_satorras import (
            EGNNLayer, SartorrasEGNN)
        from point_vs.models.geometric.egnn_multitask import (
            MultitaskSatorrasEGNN)
        from point_vs.models.geometric.pnn_geometric_base import PygLinearPass
        from point_vs.preprocessing.preprocessing import generate_edges
    ns = types.SimpleNamespace(
        EGNNLayer=EGNNLayer, SartorrasEGNN=SartorrasEGNN,
        MultitaskSatorrasEGNN=MultitaskSatorrasEGNN,
        PygLinearPass=PygLinearPass, generate_edges=generate_edges,
        Data=_Data, collate=_collate)
    return ns
