"""BASELINE configs[4]: ligand poses against one resident pocket, poses/s.

The ~800-atom synthetic pocket is uploaded once; each step ships only the
ligand atoms of 128 poses (30 atoms each), K0 assembles the complexes on the
device (crop radius `--radius`; the default keeps the whole pocket, which is
configs[4] as `synthetic_pocket_poses` defines it), K1 builds the graphs and
the 8 x 64 `egnn` scores them; scores return through pinned buffers.
Prints one JSON line.
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
import pointvs_b200 as pv  # noqa: E402
from pointvs_b200 import data  # noqa: E402
from pointvs_b200.graph import radius_graph_batch  # noqa: E402
from pointvs_b200.synthetic import synthetic_pocket_poses  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--poses', type=int, default=128 * 60)
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--pool', type=int, default=1024)
    ap.add_argument('--radius', type=float, default=1e9)
    ap.add_argument('--math', default='bf16x3')
    args = ap.parse_args()
    dev = torch.device('cuda')
    n_lig, n_pocket, n_types = 30, 800, 12
    coords, bp, feats, cptr = synthetic_pocket_poses(0, args.pool, n_pocket,
                                                     n_lig)
    types = feats[:, :n_types].argmax(1)
    per = n_lig + n_pocket
    pocket = (torch.from_numpy(coords[n_lig:per].copy()).to(dev),
              torch.ones(n_pocket, dtype=torch.uint8, device=dev),
              torch.from_numpy((types[n_lig:per] + n_types).astype(np.int16)
                               ).to(dev))
    torch.cuda.synchronize()   # the pocket is resident before K0's side stream reads it
    pool = [data.Ligand(coords[p * per:p * per + n_lig].copy(),
                        np.ones(n_lig, dtype=np.uint8),
                        types[p * per:p * per + n_lig].astype(np.int16))
            for p in range(args.pool)]
    torch.manual_seed(0)
    model = pv.SartorrasEGNN(Path('/tmp/pvs_bench'), 0, 0, None, None,
                             silent=True, **bench.MODEL_KW).to(dev).eval()
    model.set_math(args.math)
    model.set_record_side_channels(False)
    model.record_embed_coords = False
    zeros = np.zeros(args.batch, dtype=np.int32)

    def step(first):
        ligs = [pool[(first + i) % args.pool] for i in range(args.batch)]
        c, b, f, cp = data.crop_batch(ligs, [pocket], zeros, args.radius,
                                      n_types, True, dev)
        csr = radius_graph_batch(c, b, cp, 4.0, 4.0, device=dev,
                                 edge_capacity='auto')
        batch = pv.PackedBatch(f, c.float(), csr, csr.complex_ptr)
        with torch.no_grad():
            return model(batch), csr, int(cp[-1])

    # parity of the device-assembled complexes with the host generator
    c, b, f, cp = data.crop_batch(pool[:4], [pocket], zeros[:4], 1e9, n_types,
                                  True, dev)
    assert np.array_equal(c.cpu().numpy(), coords[:4 * per])
    assert np.array_equal(f.cpu().numpy(), feats[:4 * per])
    steps = args.poses // args.batch
    for i in range(12):
        step(i * args.batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    atoms, edges = 0, torch.zeros(1, dtype=torch.int64, device=dev)
    outs = torch.empty((steps, args.batch, 1), dtype=torch.float32).pin_memory()
    e0.record()
    for i in range(steps):
        out, csr, n = step(i * args.batch)
        outs[i].copy_(out.reshape(args.batch, 1), non_blocking=True)
        atoms += n
        edges += csr.n_edges_dev
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    n_scored = steps * args.batch
    print(json.dumps({
        'metric': 'ligand poses scored per second against one resident pocket '
                  '(BASELINE configs[4] shape)',
        'value': round(n_scored / (ms * 1e-3), 1), 'unit': 'poses/s',
        'poses': n_scored, 'ms_per_step': round(ms / steps, 3),
        'batch': args.batch, 'math': args.math,
        'crop_radius': None if args.radius > 1e8 else args.radius,
        'atoms_per_complex': round(atoms / n_scored, 1),
        'edges_per_pose': round(int(edges.item()) / n_scored, 1),
        'h2d_bytes_per_pose': n_lig * (24 + 1 + 2),
        'scores_finite': bool(torch.isfinite(outs).all())}))


if __name__ == '__main__':
    main()
