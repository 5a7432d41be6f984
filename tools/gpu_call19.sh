#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
for v in default a3b3 a4b2; do for k in batch c4; do
  if [ $v = default ]; then timeout 120 python tools/ab_tc.py $k default; else SLDM_LIB_PATH=build/ab/$v.so timeout 120 python tools/ab_tc.py $k $v; fi
done; done 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/ab_tc_19.jsonl
