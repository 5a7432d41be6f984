#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
for v in default a3b3 a4b2; do
  for k in batch; do
    if [ $v = default ]; then timeout 120 python tools/ab_tc.py $k default; else SLDM_LIB_PATH=build/ab/$v.so timeout 120 python tools/ab_tc.py $k $v; fi
  done
done 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/ab_tc_4.jsonl
SLDM_TC_TRACE=gpurun_out/trace_fwd_r02c.txt timeout 120 python tools/prof_kernels.py fwd > /dev/null 2>&1
SLDM_LIB_PATH=build/ab/a3b3.so SLDM_TC_TRACE=gpurun_out/trace_fwd_r02c_a3b3.txt timeout 120 python tools/prof_kernels.py fwd > /dev/null 2>&1
