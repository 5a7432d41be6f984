#!/bin/bash
# round 2, call 32: nanosleep between polls of the drain warps' mbarrier waits in k_sage_tc (dsA_B: A ns for the accumulator wait, B ns for the hand-off tile waits)
set -uo pipefail
mkdir -p gpurun_out
out=gpurun_out/ab_tc_32.jsonl; : > $out
for rep in 1 2; do
for v in default ds20_0 ds50_0 ds100_0 ds200_0 ds50_50 ds100_100; do
  if [ $v = default ]; then timeout 120 python tools/ab_tc.py batch default; else SLDM_LIB_PATH=build/ab/$v.so timeout 120 python tools/ab_tc.py batch $v; fi
done; done 2>&1 | grep -E "^\{|Error|error" | tee -a $out | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['tag'], d['proj_fwd_train'], d['proj_fwd_infer'], d['dgrad'], d['wgrad'], d['hash_fwd'], d['hash_bwd'])"
