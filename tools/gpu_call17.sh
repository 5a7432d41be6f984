#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
timeout 120 python tools/ab_tc.py batch split_stats 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/ab_tc_17.jsonl
timeout 120 python tools/ab_tc.py c4 split_stats 2>&1 | grep -E "^\{|Error|error" | tee -a gpurun_out/ab_tc_17.jsonl
SLDM_TC_TRACE=gpurun_out/trace_fwd_r02i.txt timeout 120 python tools/prof_kernels.py fwd > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_tc_paths_gpu.py tests/test_bf16_gpu.py tests/test_bench_shapes_gpu.py tests/test_properties_gpu.py -x -q -m gpu 2>&1 | tail -4
