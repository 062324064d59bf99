#!/usr/bin/env python
"""Build a variant of the C-ABI library for A/B timing (scripts/ab_bench.py):

    python scripts/build_variant.py NAME [-DFLAG ...] [--rev GITREV]
                                    [--swap FILE.cu=ALT_PATH | FILE.cu=GITREV:]

writes pointvs_b200/_C/variants/libpvs_NAME.so (git-ignored, travels to the
GPU box).  --rev builds the csrc/ tree of another commit (e.g. the previous
round's kernel as the baseline of the comparison); --swap replaces ONE source
of the current tree (by another file, or by that file at a git revision when
the value ends with ':'), which keeps the C-ABI structs current."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointvs_b200 import build as B  # noqa: E402


def main():
    name = sys.argv[1]
    args = sys.argv[2:]
    rev = None
    if '--rev' in args:
        i = args.index('--rev')
        rev = args[i + 1]
        del args[i:i + 2]
    swaps = {}
    while '--swap' in args:
        i = args.index('--swap')
        k, v = args[i + 1].split('=', 1)
        swaps[k] = v
        del args[i:i + 2]
    out_dir = os.path.join(B.OUT_DIR, 'variants')
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f'libpvs_{name}.so')
    csrc, inc = B.CSRC, os.path.join(ROOT, 'include')
    tmp = None
    if rev:
        tmp = tempfile.mkdtemp(prefix='pvs_rev_')
        subprocess.run(f'git -C {ROOT} archive {rev} pointvs_b200/csrc include | tar -x -C {tmp}',
                       shell=True, check=True)
        csrc, inc = os.path.join(tmp, 'pointvs_b200/csrc'), os.path.join(tmp, 'include')
    srcs = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith('.cu'))
    for k, v in swaps.items():
        alt = v
        if v.endswith(':'):
            alt = os.path.join(tempfile.mkdtemp(prefix='pvs_swap_'), k)
            with open(alt, 'w') as fh:
                subprocess.run(['git', '-C', ROOT, 'show', f'{v}pointvs_b200/csrc/{k}'],
                               stdout=fh, check=True)
        srcs = [alt if os.path.basename(s) == k else s for s in srcs]
    flags = [f for f in B.NVCC_FLAGS if not f.startswith('--use_fast_math')]
    cmd = ['nvcc'] + flags + args + ['-I', inc, '-I', csrc, '-o', out] + srcs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode:
        raise SystemExit(res.stdout + res.stderr)
    print(out)


if __name__ == '__main__':
    main()
