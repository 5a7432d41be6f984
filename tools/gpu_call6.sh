#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
for k in batch c4; do timeout 120 python tools/ab_tc.py $k default; done 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/ab_tc_6.jsonl
SLDM_TC_TRACE=gpurun_out/trace_fwd_r02e.txt timeout 120 python tools/prof_kernels.py fwd > /dev/null 2>&1
bash tools/gpu_validate.sh
