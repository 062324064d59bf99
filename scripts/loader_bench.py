"""Row N2 measurement: types file + parquets on disk -> scores, complexes/s.

Repeats the seven sample complexes under tests/golden/complexes into a types
file of `--n` lines (the page cache serves the files, so this measures parsing,
cropping, packing, H2D and the device path, not the disk) and scores them with
the 8 x 64 `egnn` through `PackedLoader`.  Prints one JSON line; beside it the
host-only rate of the loader threads.
"""
import argparse
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from pointvs_b200 import data, SartorrasEGNN  # noqa: E402
from tests.golden.loader_configs import CONFIGS, ROOT as COMPLEXES  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=2048)
    ap.add_argument('--batch_size', type=int, default=128)
    ap.add_argument('--workers', type=int, default=8)
    ap.add_argument('--math', default='bf16x3')
    ap.add_argument('--host-crop', action='store_true')
    ap.add_argument('--processes', action='store_true',
                    help='loader workers are spawned processes, not threads')
    args = ap.parse_args()
    lines = [ln for ln in (COMPLEXES / 'pose.types').read_text().splitlines()
             if ln.strip()]
    cfg = CONFIGS['smina_r10_e4']
    with tempfile.TemporaryDirectory() as tmp:
        types = Path(tmp) / 'many.types'
        types.write_text('\n'.join(lines[i % len(lines)]
                                   for i in range(args.n)) + '\n')
        dl = data.get_data_loader(
            COMPLEXES, types_fname=types, batch_size=args.batch_size,
            mode='val', rot=False, num_workers=args.workers, device='cuda',
            device_crop=not args.host_crop, worker_processes=args.processes,
            **cfg)
        ds = dl.dataset
        torch.manual_seed(0)
        model = SartorrasEGNN(
            Path(tmp), 0, 0, None, None, silent=True, dim_input=ds.feature_dim,
            dim_output=1, k=64, num_layers=8, graphnorm=False,
            edge_attention=True, node_attention=True, residual=True,
            normalize=True, tanh=True).cuda().eval()
        model.set_math(args.math)
        model.set_record_side_channels(False)
        model.record_embed_coords = False

        def run():
            outs, atoms = [], 0
            with torch.no_grad():
                for batch in dl:
                    outs.append(model(batch))
                    atoms += int(batch.x.shape[0])
            scores = torch.cat(outs).cpu()
            return scores, atoms

        run()                                   # warm-up (page cache, JIT-free)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        scores, atoms = run()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0

        # host-only: the loader workers without the device
        t0 = time.perf_counter()
        n_host = sum(len(items) for items, _ in dl.prepared())
        host = time.perf_counter() - t0
        assert n_host == args.n
        dl.close()

    print(json.dumps({
        'metric': 'complexes scored per second from types file + parquets',
        'value': round(args.n / wall, 1), 'unit': 'complexes/s',
        'n_complexes': args.n, 'atoms_per_complex': round(atoms / args.n, 1),
        'batch_size': args.batch_size, 'loader_threads': args.workers,
        'math': args.math, 'crop': 'host' if args.host_crop else 'device',
        'workers': 'processes' if args.processes else 'threads',
        'host_loader_only_complexes_per_s': round(args.n / host, 1),
        'scores_finite': bool(torch.isfinite(scores).all())}))


if __name__ == '__main__':
    main()
