#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench and the ncu launch
# list.  Logs land in gpurun_out/ (merged back into the repo's gpurun_out/).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/nvsmi.txt
# tcgen05 tests run in their own process: a trap there must not poison the rest
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q --tb=short --timeout 120 -p no:cacheprovider -s > gpurun_out/pytest_tc.log 2>&1
echo "pytest tc exit $?" >> gpurun_out/pytest_tc.log
tail -30 gpurun_out/pytest_tc.log
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q --tb=short --timeout 120 -p no:cacheprovider > gpurun_out/pytest_bwd.log 2>&1
echo "pytest bwd exit $?" >> gpurun_out/pytest_bwd.log
tail -40 gpurun_out/pytest_bwd.log
timeout 1200 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider --deselect tests/test_gpu_tc.py --deselect tests/test_gpu_backward.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
for m in ${BENCH_MATHS:-fp32 bf16x3 bf16}; do
  timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 --math $m > gpurun_out/bench_$m.log 2>&1
  echo "bench $m exit $?" >> gpurun_out/bench_$m.log
  tail -3 gpurun_out/bench_$m.log | cut -c1-400
done
if [ -n "$WITH_NCU" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --math ${NCU_MATH:-bf16x3} \
      > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit $?"
fi
