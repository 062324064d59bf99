"""Where a tile's time goes inside the tcgen05 edge kernel.

Builds the library with -DPVS_PHASE_PROF (clock64 at the phase boundaries of
each 4-warp group, summed in a device array) into
pointvs_b200/_C/libpvs_b200_prof.so, runs a few scoring passes of the bench
workload and prints the share of group-cycles per phase.  Debug tool: the
shipped library has none of this code.

    python scripts/phase_prof.py --build        # here (no GPU needed)
    python scripts/phase_prof.py --run          # on the GPU box
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
PROF_LIB = Path(os.environ.get('PVS_PROF_LIB') or ROOT / 'pointvs_b200' / '_C' / 'libpvs_b200_prof.so')
PHASES = ['0 rp+geometry', '1 gather+SiLU -> A', '2 GEMM1 wait',
          '3 epilogue 1', '4 M-reduce + m_out', '5 GEMM2 wait',
          '6 epilogue 2', '7 coord sums']


def build():
    sys.path.insert(0, str(ROOT))
    from pointvs_b200 import build as b
    flags = [f for f in b.NVCC_FLAGS if not f.startswith('--use_fast_math')]
    cmd = ['nvcc'] + flags + ['-DPVS_PHASE_PROF', '-I', str(ROOT / 'include'),
                              '-I', b.CSRC, '-o', str(PROF_LIB)] + b.sources()
    subprocess.run(cmd, check=True)
    print(PROF_LIB)


def run(math, steps):
    os.environ['PVS_B200_LIB'] = str(PROF_LIB)
    sys.path.insert(0, str(ROOT))
    import torch
    import bench
    import pointvs_b200 as pv
    from pointvs_b200 import _cabi
    from pointvs_b200.synthetic import synthetic_batch
    torch.manual_seed(0)
    model = pv.SartorrasEGNN(Path('/tmp/pvs_bench'), 0, 0, None, None,
                             silent=True, **bench.MODEL_KW).cuda().eval()
    model.set_math(math)
    model.set_record_side_channels(False)
    model.record_embed_coords = False
    batches = []
    for s_ in range(2):
        coords, bp, feats, cptr = synthetic_batch(10_000 * s_, 128, 1000, 30)
        batches.append(pv.PackedBatch.from_arrays(
            coords, bp, feats, cptr, bench.EDGE_RADIUS, bench.EDGE_RADIUS,
            device='cuda'))
    lib = _cabi.lib()
    lib.pvs_debug_phase_cycles.argtypes = [C.c_void_p, C.c_int]
    with torch.no_grad():
        for b in batches[:2]:
            model(b)
        lib.pvs_debug_phase_cycles(None, 1)
        for i in range(steps):
            model(batches[i % len(batches)])
    out = (C.c_ulonglong * 16)()
    lib.pvs_debug_phase_cycles(C.cast(out, C.c_void_p), 0)
    cyc = [int(out[i]) for i in range(8)]
    tot = sum(cyc)
    res = {PHASES[i]: round(cyc[i] / tot, 4) for i in range(8)}
    res['total_group_cycles_per_step'] = tot // steps
    print(json.dumps(res, indent=1))
    if hasattr(lib, 'pvs_debug_node_phase_cycles'):
        lib.pvs_debug_node_phase_cycles.argtypes = [C.c_void_p, C.c_int]
        lib.pvs_debug_node_phase_cycles(C.cast(out, C.c_void_p), 0)
        names = ['0 prologue (weights, TMEM)', '1 h block -> A', '2 GEMM 1a wait',
                 '3 M block -> A', '4 GEMM 1b wait', '5 epilogue 1 (SiLU -> A)',
                 '6 GEMM 2 wait', '7 epilogue 2 (-> staging)', '8 store pass']
        cyc = [int(out[i]) for i in range(9)]
        tot = sum(cyc)
        res = {names[i]: round(cyc[i] / tot, 4) for i in range(9)}
        # cumulative over (warm-up + steps) passes; shares are what matters
        res['group_cycles_per_launch'] = tot // ((steps + 2) * 8 * 148 * 5)
        print('node_tc_kernel (all passes since load):')
        print(json.dumps(res, indent=1))


BWD_PHASES = ['0 tile setup (row_ptr window)', '1 S0 geometry', '2 S1 gather + SiLU -> S1/SG',
              '3 G1 wait', '4 E1 m -> M tile', '5 G2 wait', '6 E2a q, craw',
              '7 E2b dcraw, dp -> X tile', '8 G3+G4 wait', '9 E3 attention bwd, dt2 -> X tile',
              '10 G5+G6 wait', '11 E4 dt1 -> SG, DT1', '12 S6a-c dd, d w_r/T, dP', '13 S6d dx rows']


def run_bwd(math, steps):
    """Phase shares of the tcgen05 edge BACKWARD kernel over training steps of
    the 16-complex batch (scripts/train_bench.py's workload)."""
    os.environ['PVS_B200_LIB'] = str(PROF_LIB)
    sys.path.insert(0, str(ROOT))
    import torch
    import pointvs_b200 as pv
    from pointvs_b200 import _cabi
    from pointvs_b200.synthetic import synthetic_batch
    kw = dict(dim_input=13, dim_output=1, k=64, num_layers=8, edge_attention=True,
              node_attention=True, residual=True, normalize=True, tanh=True,
              graphnorm=False, model_task='classification')
    model = pv.MultitaskSatorrasEGNN(Path('/tmp/pvs_train'), 1e-3, 1e-4, None, None,
                                     silent=True, **kw).cuda().train()
    model.set_math(math)
    model.set_record_side_channels(False)
    coords, bp, feats, cptr = synthetic_batch(0, 16, 1000, 30)
    y = torch.tensor([i % 2 for i in range(16)], dtype=torch.float32, device='cuda')
    c, b, f = (torch.from_numpy(a_).cuda() for a_ in (coords, bp, feats))
    lib = _cabi.lib()
    lib.pvs_debug_bwd_phase_cycles.argtypes = [C.c_void_p, C.c_int]

    def step():
        batch = pv.PackedBatch.from_arrays(c, b, f, cptr, 4.0, 4.0, y=y, device='cuda',
                                           edge_capacity='auto')
        batch.lig_fname = batch.rec_fname = [''] * 16
        yp, yt, _, _ = model.unpack_input_data_and_predict(batch)
        model.backprop(yt, yp, sync=False)
    for _ in range(2):
        step()
    lib.pvs_debug_bwd_phase_cycles(None, 1)
    for _ in range(steps):
        step()
    out = (C.c_ulonglong * 16)()
    lib.pvs_debug_bwd_phase_cycles(C.cast(out, C.c_void_p), 0)
    cyc = [int(out[i]) for i in range(14)]
    tot = sum(cyc)
    res = {BWD_PHASES[i]: round(cyc[i] / tot, 4) for i in range(14)}
    res['cta_cycles_per_launch'] = tot // (steps * 8)
    res['cycles_per_cta_per_launch'] = tot // (steps * 8 * 148)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--build', action='store_true')
    ap.add_argument('--run', action='store_true')
    ap.add_argument('--bwd', action='store_true',
                    help='with --run: the edge backward kernel over training steps')
    ap.add_argument('--math', default='bf16x3')
    ap.add_argument('--steps', type=int, default=3)
    a = ap.parse_args()
    if a.build:
        build()
    if a.run and a.bwd:
        run_bwd(a.math, a.steps)
    elif a.run:
        run(a.math, a.steps)
