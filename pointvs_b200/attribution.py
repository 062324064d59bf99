"""Attribution callers of the scoring path (N3).

Mirror of /root/reference/point_vs/attribution/attribution_fns.py: the same
function names, arguments and return arrays, for the EGNN family only
(`isinstance(model, PNNGeometricBase)` branches of the reference).

Atom masking (:365-442) and bond masking (:35-109) score N (or E) copies of
one graph with one atom (or the two atoms of a ligand-receptor edge) removed.
The reference runs them one forward at a time; here the masked copies are
built on the device from the caller's edge list (no distance search: masking
only deletes edges, it does not re-derive them) and scored `bs` copies per
launch through the arbitrary-order edge path.  Models with GraphNorm are
scored one copy at a time, because the reference's GraphNorm statistics run
over whatever is in the batch (egnn_satorras.py:84).

The attention / coordinate-tracking attributions read the per-layer side
channels after one forward, as the reference does.
"""
from types import SimpleNamespace

import numpy as np
import torch
from scipy.stats import rankdata

from .dense import run_head
from .egnn import EGNNLayer, PNNGeometricBase

# The reference toggles this module global from attribute.py (:17).
SIGMOID = False


def attention_wrapper(**kwargs):
    """Dummy fn (reference :21-23)."""


def cam_wrapper(**kwargs):
    """Dummy fn (reference :26-28)."""


def masking_wrapper(**kwargs):
    """Dummy fn (reference :31-32)."""


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
def _device(model):
    return next(model.parameters()).device


def _single_graph(model, p, v, edge_indices, edge_attrs):
    """`get_pyg_single_graph_for_inference(Data(...))`: a batch of one."""
    dev = _device(model)
    x = torch.as_tensor(v).squeeze().to(dev)
    pos = torch.as_tensor(p).squeeze().to(dev).float().clone()
    if x.dim() == 1:
        x = x.unsqueeze(0)
    if pos.dim() == 1:
        pos = pos.unsqueeze(0)
    n = x.shape[0]
    return SimpleNamespace(
        x=x, pos=pos, edge_index=torch.as_tensor(edge_indices).to(dev).long(),
        edge_attr=torch.as_tensor(edge_attrs).to(dev),
        batch=torch.zeros(n, dtype=torch.long, device=dev), y=None,
        lig_fname=[None], rec_fname=[None])


def _uses_graphnorm(model):
    return any(isinstance(l, EGNNLayer) and l.graphnorm
               for l in model.layers)


def _forward_with_side_channels(model, graph):
    saved = [l.record_side_channels for l in model.layers
             if isinstance(l, EGNNLayer)]
    embed_saved = getattr(model, 'record_embed_coords', True)
    model.set_record_side_channels(True)
    model.record_embed_coords = True
    try:
        with torch.no_grad():
            out = model(graph)
    finally:
        for layer, on in zip([l for l in model.layers
                              if isinstance(l, EGNNLayer)], saved):
            layer.record_side_channels = on
        model.record_embed_coords = embed_saved
    return out


def _score_rows(out, n_graphs, sigmoid, mode='atom'):
    """Model output of a batch -> one float per graph.  'atom': the
    reference's atom_masking bookkeeping (:392-400), the mean of a 3-wide
    regression output, else the (optionally sigmoided) single logit.  'bond':
    `extract_score` of bond_masking (:38-46), element 1 of a multi-output
    score after the optional sigmoid."""
    out = out.reshape(n_graphs, -1)
    if mode == 'atom':
        if out.shape[1] == 3:
            return out.mean(dim=1)
        if out.shape[1] != 1:
            raise ValueError('atom masking expects 1 or 3 outputs per complex')
        out = out[:, 0]
    else:
        out = out[:, 1] if out.shape[1] > 1 else out[:, 0]
    return torch.sigmoid(out) if sigmoid else out


def masked_copies(x, pos, edge_index, edge_attr, removed):
    """Pack len(removed) copies of one graph, copy c without the atoms in
    `removed[c]` (LongTensor [V, R], -1 = nothing), renumbering the survivors
    and dropping every edge that touches a removed atom.  All on the device.

    Returns a graph namespace (x, pos, edge_index, edge_attr, batch) whose
    copies appear in order, each with the original node and edge order."""
    dev = x.device
    n, e = x.shape[0], edge_index.shape[1]
    removed = removed.to(dev)
    v = removed.shape[0]
    keep = torch.ones((v, n), dtype=torch.bool, device=dev)
    rows = torch.arange(v, device=dev).unsqueeze(1).expand_as(removed)
    valid = removed >= 0
    keep[rows[valid], removed[valid]] = False
    counts = keep.sum(dim=1)
    offsets = torch.cumsum(counts, 0) - counts
    new_index = torch.cumsum(keep, dim=1) - 1 + offsets.unsqueeze(1)
    copy_of, node_of = torch.nonzero(keep, as_tuple=True)
    src, dst = edge_index[0], edge_index[1]
    ekeep = keep[:, src] & keep[:, dst]                       # [V, E]
    ecopy, eid = torch.nonzero(ekeep, as_tuple=True)
    del e
    return SimpleNamespace(
        x=x[node_of], pos=pos[node_of].clone(),
        edge_index=torch.stack([new_index[ecopy, src[eid]],
                                new_index[ecopy, dst[eid]]]),
        edge_attr=edge_attr[eid], batch=copy_of, y=None,
        lig_fname=[None] * v, rec_fname=[None] * v, num_graphs=v)


def _masked_scores(model, graph, removed, bs, sigmoid, mode='atom'):
    """Scores of the masked copies, `bs` per forward."""
    if _uses_graphnorm(model):
        bs = 1
    saved = [l.record_side_channels for l in model.layers
             if isinstance(l, EGNNLayer)]
    embed_saved = getattr(model, 'record_embed_coords', True)
    model.set_record_side_channels(False)
    model.record_embed_coords = False
    outs = []
    try:
        with torch.no_grad():
            for start in range(0, removed.shape[0], max(1, bs)):
                chunk = removed[start:start + max(1, bs)]
                batch = masked_copies(graph.x, graph.pos, graph.edge_index,
                                      graph.edge_attr, chunk)
                if batch.x.shape[0] == 0:
                    raise ValueError('masking removed every atom')
                outs.append(_score_rows(model(batch), chunk.shape[0], sigmoid,
                                        mode))
    finally:
        for layer, on in zip([l for l in model.layers
                              if isinstance(l, EGNNLayer)], saved):
            layer.record_side_channels = on
        model.record_embed_coords = embed_saved
    return torch.cat(outs) if outs else torch.zeros(0, device=graph.x.device)


def _check(model):
    if not isinstance(model, PNNGeometricBase):
        raise TypeError('attribution here covers the EGNN family only '
                        '(PNNGeometricBase subclasses)')


# --------------------------------------------------------------------------
# masking
# --------------------------------------------------------------------------
def atom_masking(model, p, v, m=None, bs=32, edge_indices=None,
                 edge_attrs=None, resis=None, **kwargs):
    """Score drop when each atom is deleted: original - masked, [n_atoms]
    (reference :365-442)."""
    del m, resis
    _check(model)
    sigmoid = kwargs.get('sigmoid', SIGMOID)
    graph = _single_graph(model, p, v, edge_indices, edge_attrs)
    n = graph.x.shape[0]
    with torch.no_grad():
        original = _score_rows(model(_clone(graph)), 1, sigmoid)
    removed = torch.arange(n, device=graph.x.device).unsqueeze(1)
    masked = _masked_scores(model, graph, removed, bs, sigmoid)
    return (original - masked).double().cpu().numpy()


def bond_masking(model, p, v, m=None, bs=32, edge_indices=None,
                 edge_attrs=None, **kwargs):
    """Score drop when both atoms of each ligand-receptor edge are deleted;
    0 for the other edges; [n_edges] (reference :35-109)."""
    del m
    _check(model)
    sigmoid = kwargs.get('sigmoid', SIGMOID)
    graph = _single_graph(model, p, v, edge_indices, edge_attrs)
    with torch.no_grad():
        original = _score_rows(model(_clone(graph)), 1, sigmoid, 'bond')
    inter = graph.edge_attr[:, 1] != 0
    idx = torch.nonzero(inter).reshape(-1)
    scores = torch.zeros(graph.edge_index.shape[1], dtype=torch.float64,
                         device=graph.x.device)
    if idx.numel():
        ends = graph.edge_index[:, idx]
        lo, hi = ends.min(dim=0).values, ends.max(dim=0).values
        removed = torch.stack([lo, hi], dim=1)
        masked = _masked_scores(model, graph, removed, bs, sigmoid, 'bond')
        scores[idx] = (original - masked).double()
    return scores.cpu().numpy()


def _clone(graph):
    return SimpleNamespace(**{**graph.__dict__, 'pos': graph.pos.clone()})


# --------------------------------------------------------------------------
# attention and coordinate tracking (side channels after one forward)
# --------------------------------------------------------------------------
def edge_attention(model, p, v, edge_indices=None, edge_attrs=None,
                   gnn_layer=-1, **kwargs):
    """Edge attention weights of one layer, caller's edge order (:295-309)."""
    _check(model)
    _forward_with_side_channels(
        model, _single_graph(model, p, v, edge_indices, edge_attrs))
    return model.layers[gnn_layer].att_val.reshape((-1,))


def node_attention(model, p, v, edge_indices=None, edge_attrs=None,
                   gnn_layer=-1, **kwargs):
    """Node attention weights of one layer (logit-transformed when SIGMOID),
    reference :259-292."""
    _check(model)
    _forward_with_side_channels(
        model, _single_graph(model, p, v, edge_indices, edge_attrs))
    att = model.layers[gnn_layer].node_att_val.reshape((-1,))
    if not kwargs.get('sigmoid', SIGMOID):
        return att
    return np.log(att / (1 - att))


def _mean_rank(model, graph, attr):
    _forward_with_side_channels(model, graph)
    ranks = []
    for idx, layer in enumerate(model.layers):
        val = getattr(layer, attr, None)
        if val is not None:
            if idx == 10:
                break
            ranks.append(rankdata(val.flatten()) - 1)
    return np.mean(np.vstack(ranks).T, axis=1)


def mean_edge_attention_rank(model, p, v, edge_indices=None, edge_attrs=None,
                             gnn_layer=-1, **kwargs):
    """Mean over layers of the rank of each edge's attention (:234-256)."""
    _check(model)
    return _mean_rank(model, _single_graph(model, p, v, edge_indices,
                                           edge_attrs), 'att_val')


def mean_node_attention_rank(model, p, v, edge_indices=None, edge_attrs=None,
                             gnn_layer=-1, **kwargs):
    """Mean over layers of the rank of each atom's attention (:212-231)."""
    _check(model)
    return _mean_rank(model, _single_graph(model, p, v, edge_indices,
                                           edge_attrs), 'node_att_val')


def track_position_changes(model, p, v, edge_indices=None, edge_attrs=None,
                           **kwargs):
    """Sum over layers of each atom's displacement from its input position
    (:136-156)."""
    _check(model)
    _forward_with_side_channels(
        model, _single_graph(model, p, v, edge_indices, edge_attrs))
    original = model.layers[0].intermediate_coords
    disp = []
    for layer in range(1, model.n_layers + 1):
        d = model.layers[layer].intermediate_coords - original
        disp.append(np.sqrt(np.sum(d ** 2, axis=1)))
    return np.sum(np.vstack(disp).T, axis=1)


def track_bond_lengths(model, p, v, edge_indices=None, edge_attrs=None,
                       **kwargs):
    """Change of every edge's length between the input and the last layer's
    coordinates (:112-133)."""
    _check(model)
    _forward_with_side_channels(
        model, _single_graph(model, p, v, edge_indices, edge_attrs))
    ei = torch.as_tensor(edge_indices).cpu().numpy()
    lengths = []
    for coords in (model.layers[0].intermediate_coords,
                   model.layers[-1].intermediate_coords):
        lengths.append(np.linalg.norm(coords[ei[0]] - coords[ei[1]], axis=1))
    return lengths[1] - lengths[0]


def cam(model, p, v, m=None, edge_indices=None, edge_attrs=None, **kwargs):
    """Class activation map: the scoring head applied to every atom's final
    embedding instead of the pooled one (:312-338)."""
    del m
    _check(model)
    graph = _single_graph(model, p, v, edge_indices, edge_attrs)
    with torch.no_grad():
        feats, edges, coords, edge_attributes, batch = model.unpack_graph(
            graph)
        feats, _ = model.get_embeddings(feats, edges, coords, edge_attributes,
                                        batch)
        x = run_head(model.feats_linear_layers, feats).cpu().numpy()
    if x.ndim == 2 and x.shape[1] == 3:
        x = np.mean(x, axis=1)
    return x


__all__ = ['atom_masking', 'bond_masking', 'edge_attention', 'node_attention',
           'mean_edge_attention_rank', 'mean_node_attention_rank',
           'track_position_changes', 'track_bond_lengths', 'cam',
           'masked_copies', 'attention_wrapper', 'cam_wrapper',
           'masking_wrapper']
