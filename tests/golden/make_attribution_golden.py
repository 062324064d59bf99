"""Generate the attribution goldens (row N3) from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_attribution_golden.py

Runs the reference's own `point_vs.attribution.attribution_fns` (imported
through oracle/ref_shim.py) with seeded reference models on one complex of
tests/golden/loader.npz and stores inputs, weights and every returned array in
tests/golden/attribution.npz.
"""
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))

from oracle import ref_shim  # noqa: E402

REF = ref_shim.import_reference()
from point_vs.attribution import attribution_fns as A  # noqa: E402

from tests.golden.make_attribution_cfg import MODELS  # noqa: E402
# track_position_changes / track_bond_lengths are left out: on a CPU run of the
# reference every layer's `intermediate_coords` aliases the one coordinate
# buffer that is updated in place (to_numpy of a CPU tensor shares memory), so
# both return zeros; tests check them against the oracle's per-layer trace.
FNS = ['atom_masking', 'bond_masking', 'edge_attention', 'node_attention',
       'mean_edge_attention_rank', 'mean_node_attention_rank', 'cam']


def main():
    gold = np.load(HERE / 'loader.npz')
    cfg, item = 'atomic_h_r6_e3', 0
    x = torch.from_numpy(gold[f'{cfg}/{item}/x'].astype(np.float32))
    pos = torch.from_numpy(gold[f'{cfg}/{item}/pos'])
    ei = torch.from_numpy(gold[f'{cfg}/{item}/edge_index'].astype(np.int64))
    ea = torch.nn.functional.one_hot(
        torch.from_numpy(gold[f'{cfg}/{item}/edge_attr'].astype(np.int64)), 3)
    out = {'x': x.numpy(), 'pos': pos.numpy(),
           'edge_index': ei.numpy().astype(np.int32),
           'edge_attr': ea.numpy().astype(np.uint8)}
    for name, kw in MODELS.items():
        torch.manual_seed(5)
        with tempfile.TemporaryDirectory() as tmp:
            model = REF.SartorrasEGNN(Path(tmp), 0, 0, None, None, silent=True,
                                      **kw)
        with torch.no_grad():     # make the coordinate path visible
            for n, p in model.named_parameters():
                if n.endswith('coord_mlp.2.weight'):
                    p.mul_(300.0)
        model.eval()
        for k, v in model.state_dict().items():
            out[f'{name}/sd.{k}'] = v.numpy()
        for sig in (False, True):
            A.SIGMOID = sig
            for fn in FNS:
                if sig and fn not in ('atom_masking', 'bond_masking',
                                      'node_attention'):
                    continue
                if fn == 'bond_masking' and (sig or kw['dim_output'] == 1):
                    # the reference's bond_masking raises TypeError for
                    # single-output models (len() of a 0-d score, :43)
                    continue
                if fn == 'atom_masking' and kw['dim_output'] != 1:
                    # ... and its atom_masking for multi-output ones (the
                    # batch-of-one output is 1-D, so `is_regression` is False)
                    continue
                if fn.endswith('node_attention') or fn.startswith('mean_node'):
                    if not kw['node_attention']:
                        continue
                with torch.no_grad():
                    res = getattr(A, fn)(
                        model, pos.clone().unsqueeze(0), x.clone().unsqueeze(0),
                        edge_indices=ei.clone(), edge_attrs=ea.clone())
                res = np.asarray(res, dtype=np.float64)
                out[f'{name}/{fn}/sigmoid{int(sig)}'] = res
                print(name, fn, sig, res.shape, float(np.abs(res).max()))
    np.savez_compressed(HERE / 'attribution.npz', **out)
    print('attribution.npz', (HERE / 'attribution.npz').stat().st_size)


if __name__ == '__main__':
    main()
