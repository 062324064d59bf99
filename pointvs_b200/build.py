"""Build the C-ABI CUDA library in-tree (pointvs_b200/_C/libpvs_b200.so).

`nvcc -gencode arch=compute_100a,code=sm_100a` cross-compiles without a GPU.
The library links cudart statically and has no libtorch dependency: the ABI
is plain C (include/pvs_b200.h).
"""
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
OUT_DIR = os.path.join(HERE, '_C')
LIB_PATH = os.path.join(OUT_DIR, 'libpvs_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
    '-std=c++17', '-Xcompiler', '-fPIC', '-shared', '-cudart', 'static',
    '--use_fast_math=false',
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + \
        glob.glob(os.path.join(ROOT, 'include', '*.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a into one shared library."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    os.makedirs(OUT_DIR, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if not f.startswith('--use_fast_math')]
    cmd = [nvcc] + flags + ['-I', os.path.join(ROOT, 'include'), '-I', CSRC,
                            '-o', LIB_PATH] + sources()
    if verbose:
        print(' '.join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force=True, verbose=True))
