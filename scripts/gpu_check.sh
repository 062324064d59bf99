#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench and the ncu launch
# list.  Logs land in gpurun_out/ (merged back into the repo's gpurun_out/).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/nvsmi.txt
timeout 1200 python -m pytest tests -m gpu -q --tb=short --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
tail -5 gpurun_out/bench.log
if [ -n "$WITH_NCU" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
      > gpurun_out/ncu_bench.log 2>&1
  echo "ncu exit $?"
fi
