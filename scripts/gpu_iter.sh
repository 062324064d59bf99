#!/bin/bash
# Quick iteration on the GPU box: tensor-core parity tests, phase profile, short bench.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tc.py tests/test_gpu_egnn.py -x -q 2>&1 | tail -3
if [ -f pointvs_b200/_C/libpvs_b200_prof.so ]; then
  timeout 300 python scripts/phase_prof.py --run 2>&1 | tail -11 | tr -d '\n'; echo
fi
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/iter_bench.log 2>&1
tail -1 gpurun_out/iter_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'edge_us',d['config'].get('edge_kernel_us_per_launch'),'roof',d['roofline']['frac'])
print({k:v for k,v in d['config'].items() if 'stage' in k or 'ms' in k})"
