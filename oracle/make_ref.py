"""TEST INFRASTRUCTURE ONLY -- recipe for oracle/_ref (the real reference, CPU).

PointVS is pure Python: "building" the reference means making its own package
importable where /root/reference does not exist (the GPU box).  This script
copies the *.py files of /root/reference/point_vs, unmodified, into
oracle/_ref/point_vs.  oracle/_ref/ is git-ignored (never part of the history)
but travels with the gpurun snapshot, like the built .so files.

    python oracle/make_ref.py          # needs /root/reference; idempotent

Users: bench.py's CPU legs (`--impl reference`, `cpu_baseline`), through
oracle/ref_shim.py, which stubs the reference's absent third-party imports.
Nothing under pointvs_b200/ may import oracle/_ref.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/point_vs'
DST = os.path.join(HERE, '_ref', 'point_vs')


def make_ref(verbose=False):
    if not os.path.isdir(SRC):
        return False
    n = 0
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d != '__pycache__']
        rel = os.path.relpath(root, SRC)
        out_dir = os.path.join(DST, rel) if rel != '.' else DST
        for f in files:
            if not f.endswith('.py'):
                continue
            os.makedirs(out_dir, exist_ok=True)
            s, d = os.path.join(root, f), os.path.join(out_dir, f)
            if (not os.path.exists(d) or
                    os.path.getmtime(d) < os.path.getmtime(s) or
                    os.path.getsize(d) != os.path.getsize(s)):
                shutil.copyfile(s, d)
            n += 1
    if verbose:
        print(f'oracle/_ref: {n} reference files under {DST}')
    return True


if __name__ == '__main__':
    ok = make_ref(verbose=True)
    if not ok:
        print('reference tree not present at /root/reference; nothing done')
    sys.exit(0)
