#!/bin/bash
mkdir -p gpurun_out
for b in 16 64; do
  timeout 300 python scripts/train_bench.py --steps 10 --batch $b > gpurun_out/train_b$b.log 2>&1
  tail -1 gpurun_out/train_b$b.log | cut -c1-300
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/train_launches.csv python scripts/train_bench.py --steps 1 --warmup 1 \
    > gpurun_out/ncu_train.log 2>&1
echo "ncu train exit $?"
