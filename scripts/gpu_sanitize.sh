#!/bin/bash
# compute-sanitizer over one small K1 + K2 (tcgen05 and FFMA) + K3 pass.
# Logs: gpurun_out/sanitize_<tool>_<math>.log (copy the summaries to profiles/).
mkdir -p gpurun_out
TOOLS=${SAN_TOOLS:-memcheck racecheck synccheck}
MATHS=${SAN_MATHS:-bf16x3 fp32}
for tool in $TOOLS; do
  for m in $MATHS; do
    extra=""
    [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
    timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --print-limit 40 \
        python scripts/sanitize_pass.py --math $m ${SAN_ARGS} \
        > gpurun_out/sanitize_${tool}_${m}.log 2>&1
    echo "sanitize $tool $m exit $?" | tee -a gpurun_out/sanitize_${tool}_${m}.log
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|forward ok|backward ok|backprop ok|Error|hazard" gpurun_out/sanitize_${tool}_${m}.log | head -8
  done
done
