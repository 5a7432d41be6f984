#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_tc_paths_gpu.py -x -q -m gpu > gpurun_out/call3_tests_tc.log 2>&1; echo "tc tests rc=$?"; tail -3 gpurun_out/call3_tests_tc.log
for v in default r1tc; do
  for k in batch c4; do
    if [ $v = default ]; then timeout 120 python tools/ab_tc.py $k default; else SLDM_LIB_PATH=build/ab/$v.so timeout 120 python tools/ab_tc.py $k $v; fi
  done
done 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/ab_tc_3.jsonl
SLDM_TC_TRACE=gpurun_out/trace_fwd_r02b.txt timeout 120 python tools/prof_kernels.py fwd > /dev/null 2>&1
