"""The optional bf16x3 edge-kernel variant `tc8` (csrc/egnn_edge_tc8.cu:
8-warp groups, message segment-reduce as an MN-major tcgen05 GEMM) is selected
per process by PVS_EDGE_TC8=1, so its parity run is a child process: the
tcgen05 layer / model parity tests of test_gpu_tc.py with the variant on."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.gpu
def test_tc8_variant_passes_the_tcgen05_parity_tests():
    env = dict(os.environ, PVS_EDGE_TC8='1')
    res = subprocess.run(
        [sys.executable, '-m', 'pytest', 'tests/test_gpu_tc.py', '-m', 'gpu',
         '-x', '-q', '-p', 'no:cacheprovider', '-W', 'ignore'],
        cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-1000:]
    assert ' passed' in res.stdout
