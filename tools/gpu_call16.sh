#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
for v in default susp100 susp1k susp10k; do
  if [ $v = default ]; then timeout 120 python tools/ab_tc.py batch default; else SLDM_LIB_PATH=build/ab/$v.so timeout 120 python tools/ab_tc.py batch $v; fi
done 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/ab_tc_16.jsonl
