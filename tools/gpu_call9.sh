#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
for v in default finu4 nostore; do
  if [ $v = default ]; then timeout 120 python tools/ab_tc.py batch default; else SLDM_LIB_PATH=build/ab/$v.so timeout 120 python tools/ab_tc.py batch $v; fi
done 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/ab_tc_9.jsonl
SLDM_LIB_PATH=build/ab/nostore.so SLDM_TC_TRACE=gpurun_out/trace_fwd_r02g_nostore.txt timeout 120 python tools/prof_kernels.py fwd > /dev/null 2>&1
timeout 120 python tools/dbg_bf16.py 2>&1 | tail -12
bash tools/gpu_validate.sh
