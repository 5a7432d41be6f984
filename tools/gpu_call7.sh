#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
for v in default nob nostore nob_nostore; do
  if [ $v = default ]; then timeout 120 python tools/ab_tc.py batch default; else SLDM_LIB_PATH=build/ab/$v.so timeout 120 python tools/ab_tc.py batch $v; fi
done 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/ab_tc_7.jsonl
timeout 600 python -m pytest tests/test_bf16_gpu.py tests/test_graphed_gpu.py -x -q -m gpu 2>&1 | tail -15
