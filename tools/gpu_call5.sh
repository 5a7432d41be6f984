#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
for v in default r1tc; do
  for k in batch c4; do
    if [ $v = default ]; then timeout 120 python tools/ab_tc.py $k default; else SLDM_LIB_PATH=build/ab/$v.so timeout 120 python tools/ab_tc.py $k $v; fi
  done
done 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/ab_tc_5.jsonl
SLDM_TC_TRACE=gpurun_out/trace_fwd_r02d.txt timeout 120 python tools/prof_kernels.py fwd > /dev/null 2>&1
bash tools/gpu_validate.sh
