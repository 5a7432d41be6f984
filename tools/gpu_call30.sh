#!/bin/bash
# round 2, call 30: L2 cache hints in k_sage_tc (A loads evict_first = 1, weight loads evict_last = 2, bulk stores evict_first = 4, finisher stores .cs = 8)
set -uo pipefail
mkdir -p gpurun_out
out=gpurun_out/ab_tc_30.jsonl; : > $out
for rep in 1 2; do
for v in default hint8 hint1 hint2 hint3 hint4 hint11 hint15; do
  if [ $v = default ]; then timeout 120 python tools/ab_tc.py batch default; else SLDM_LIB_PATH=build/ab/$v.so timeout 120 python tools/ab_tc.py batch $v; fi
done; done 2>&1 | grep -E "^\{|Error|error" | tee -a $out | cut -c1-400
for v in default hint15; do
  if [ $v = default ]; then timeout 120 python tools/ab_tc.py c4 default; else SLDM_LIB_PATH=build/ab/$v.so timeout 120 python tools/ab_tc.py c4 $v; fi
done 2>&1 | grep -E "^\{|Error|error" | tee -a $out | cut -c1-400
