#!/bin/bash
# ncu --set full captures of the hot kernels (one GPU, short command).
mkdir -p gpurun_out
MATH=${NCU_MATH:-bf16x3}
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"${NCU_KERNEL:-egnn_edge_tc}" -s ${NCU_SKIP:-8} -c ${NCU_COUNT:-2} -f -o gpurun_out/${NCU_OUT:-prof_edge} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --math $MATH > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
if [ -n "$NCU_KERNEL2" ]; then
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"$NCU_KERNEL2" -s ${NCU_SKIP2:-0} -c ${NCU_COUNT2:-4} -f -o gpurun_out/${NCU_OUT2:-prof_other} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --math $MATH > gpurun_out/ncu_full2.log 2>&1
echo "ncu full 2 exit $?"
fi
ls -la gpurun_out/*.ncu-rep
