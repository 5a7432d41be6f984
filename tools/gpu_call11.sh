#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
timeout 120 python tools/ab_tc.py batch default 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/ab_tc_11.jsonl
timeout 120 python tools/ab_tc.py c4 default 2>&1 | grep -E "^\{|Error|error" | tee -a gpurun_out/ab_tc_11.jsonl
SLDM_TC_TRACE=gpurun_out/trace_fwd_r02h.txt timeout 120 python tools/prof_kernels.py fwd > /dev/null 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/call11_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/call11_tests.log
