#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
timeout 120 python tools/ab_tc.py batch default 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/ab_tc_8.jsonl
SLDM_TC_TRACE=gpurun_out/trace_fwd_r02f.txt timeout 120 python tools/prof_kernels.py fwd > /dev/null 2>&1
timeout 120 python tools/dbg_bf16.py 2>&1 | tail -30
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_tc_paths_gpu.py tests/test_graphed_gpu.py -x -q -m gpu 2>&1 | tail -8
