#!/bin/bash
# Multi-GPU run (gpurun --gpus N): correctness check, scaling bench, DP training.
N=${N:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/multi_check_n$N.log 2>&1
echo "multi check exit $?"; tail -3 gpurun_out/multi_check_n$N.log
timeout 600 $RUN --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --math ${MATH:-bf16x3} > gpurun_out/bench_n$N.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --math ${MATH:-bf16x3} --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1
tail -1 gpurun_out/bench_n1.log | cut -c1-200
timeout 600 $RUN --master-port 29513 scripts/train_bench.py --steps 10 > gpurun_out/train_n$N.log 2>&1
echo "train exit $?"; tail -1 gpurun_out/train_n$N.log | cut -c1-400
timeout 300 $RUN --master-port 29514 bench.py --impl reference --gpus $N --steps 2 --warmup 1 --cpu-sample 4 > gpurun_out/bench_ref_n$N.log 2>&1
echo "ref exit $?"; tail -1 gpurun_out/bench_ref_n$N.log | cut -c1-200
