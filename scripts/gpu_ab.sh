#!/bin/bash
# One gpurun call: parity tests of the tcgen05 path on the main build, then A/B
# timing of every library under pointvs_b200/_C/variants (scripts/ab_bench.py).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_egnn.py tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py -m gpu -x -q --tb=short --timeout 300 -p no:cacheprovider > gpurun_out/ab_pytest.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/ab_pytest.log
tail -5 gpurun_out/ab_pytest.log
timeout 900 python scripts/ab_bench.py ${AB_LIBS:-pointvs_b200/_C/variants/*.so} --rounds ${AB_ROUNDS:-2} --math ${AB_MATH:-bf16x3} 2>&1 | tee gpurun_out/ab_bench.log
