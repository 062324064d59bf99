"""A/B timing of library builds on the same GPU in one call.

    python scripts/ab_bench.py libA.so libB.so [--rounds 3]

Each build runs in its own subprocess (PVS_B200_LIB), interleaved A B A B ...;
the workload is the bench's scoring pass on one prebuilt packed batch (graph
construction excluded), timed with CUDA events.  Prints ms per pass per build.
"""
import argparse
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def child(math, passes):
    sys.path.insert(0, str(ROOT))
    import torch
    import bench
    import pointvs_b200 as pv
    from pointvs_b200.synthetic import synthetic_batch
    torch.manual_seed(0)
    model = pv.SartorrasEGNN(Path('/tmp/pvs_bench'), 0, 0, None, None,
                             silent=True, **bench.MODEL_KW).cuda().eval()
    model.set_math(math)
    model.set_record_side_channels(False)
    model.record_embed_coords = False
    coords, bp, feats, cptr = synthetic_batch(0, 128, 1000, 30)
    batch = pv.PackedBatch.from_arrays(coords, bp, feats, cptr,
                                       bench.EDGE_RADIUS, bench.EDGE_RADIUS,
                                       device='cuda')
    pos0 = batch.pos.clone()
    with torch.no_grad():
        for _ in range(5):
            batch.pos.copy_(pos0)
            out = model(batch)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(True), torch.cuda.Event(True)
        t0.record()
        for _ in range(passes):
            batch.pos.copy_(pos0)
            out = model(batch)
        t1.record()
        torch.cuda.synchronize()
    print(json.dumps({'ms': t0.elapsed_time(t1) / passes,
                      'checksum': float(out.double().sum())}))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('libs', nargs='*')
    ap.add_argument('--rounds', type=int, default=3)
    ap.add_argument('--math', default='bf16x3')
    ap.add_argument('--passes', type=int, default=40)
    ap.add_argument('--child', action='store_true')
    a = ap.parse_args()
    if a.child:
        child(a.math, a.passes)
        sys.exit(0)
    res = {lib: [] for lib in a.libs}
    sums = {}
    for _ in range(a.rounds):
        for lib in a.libs:
            env = dict(os.environ, PVS_B200_LIB=str(Path(lib).resolve()))
            out = subprocess.run(
                [sys.executable, __file__, '--child', '--math', a.math,
                 '--passes', str(a.passes)], env=env, capture_output=True,
                text=True)
            if out.returncode:
                print(lib, 'FAILED', out.stderr[-400:])
                continue
            d = json.loads(out.stdout.strip().splitlines()[-1])
            res[lib].append(round(d['ms'], 4))
            sums[lib] = d['checksum']
    for lib in a.libs:
        print(Path(lib).name, 'ms/pass', res[lib], 'min',
              min(res[lib]) if res[lib] else None, 'checksum', sums.get(lib))
