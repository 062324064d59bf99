"""One small K0/K1/K2/K3 pass for compute-sanitizer (memcheck / racecheck /
synccheck / initcheck).  Small on purpose: the tools slow kernels 10-100x.

    compute-sanitizer --tool racecheck python scripts/sanitize_pass.py [--math bf16x3]

Covers: K1 radius graph (count + emit + tiles), K2 forward with the tcgen05
edge/node kernels (or the FFMA kernels with --math fp32), edge-packed tile
fix-up, mean pool + head, and K3 backward (multitask model, BCE loss).
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--math', default='bf16x3')
    ap.add_argument('--complexes', type=int, default=3)
    ap.add_argument('--atoms', type=int, default=260)
    ap.add_argument('--layers', type=int, default=2)
    ap.add_argument('--k', type=int, default=64)
    ap.add_argument('--no-backward', action='store_true')
    a = ap.parse_args()
    import torch
    import pointvs_b200 as pv
    from pointvs_b200.synthetic import synthetic_batch

    kw = dict(dim_input=13, dim_output=1, k=a.k, num_layers=a.layers,
              edge_attention=True, node_attention=True, residual=True,
              normalize=True, tanh=True, graphnorm=False)
    torch.manual_seed(0)
    model = pv.MultitaskSatorrasEGNN(Path('/tmp/pvs_sanitize'), 0, 0, None, None,
                                     silent=True, **kw).cuda()
    model.set_math(a.math)
    coords, bp, feats, cptr = synthetic_batch(7, a.complexes, a.atoms, 20)
    batch = pv.PackedBatch.from_arrays(coords, bp, feats, cptr, 4.0, 4.0,
                                       device='cuda')
    model.eval()
    with torch.no_grad():
        s = model(batch)
    torch.cuda.synchronize()
    print('forward ok', s.reshape(-1).tolist())
    if not a.no_backward:
        model.train()
        batch = pv.PackedBatch.from_arrays(coords, bp, feats, cptr, 4.0, 4.0,
                                           device='cuda')
        out = model(batch).reshape(-1)
        y = torch.ones_like(out)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(out, y)
        loss.backward()
        torch.cuda.synchronize()
        gn = sum(float(p.grad.abs().sum()) for p in model.parameters()
                 if p.grad is not None)
        print('backward ok', float(loss), gn)
        # the training loop's own path: capacity-bounded graph (no host sync),
        # gradient arena, lean stacked pass from the second step on, symmetric
        # CSC, grouped weight gradients, arena clamp + fused Adam
        model2 = pv.MultitaskSatorrasEGNN(Path('/tmp/pvs_sanitize'), 1e-3, 1e-4, None,
                                          None, silent=True, model_task='classification',
                                          **kw).cuda().train()
        model2.set_math(a.math)
        model2.set_record_side_channels(False)
        ytrue = torch.tensor([float(i % 2) for i in range(a.complexes)], device='cuda')
        losses = []
        for _ in range(3):
            b2 = pv.PackedBatch.from_arrays(coords, bp, feats, cptr, 4.0, 4.0, y=ytrue,
                                            device='cuda', edge_capacity='auto')
            b2.lig_fname = b2.rec_fname = [''] * a.complexes
            yp, yt, _, _ = model2.unpack_input_data_and_predict(b2)
            losses.append(float(model2.backprop(yt, yp, sync=False)))
        torch.cuda.synchronize()
        b2.pvs_csr.check_overflow()
        print('backprop ok', losses, 'lean', getattr(model2, '_lean_training', False))


if __name__ == '__main__':
    main()
