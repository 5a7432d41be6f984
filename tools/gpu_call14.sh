#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
for l in 4 16; do for k in batch c4; do SLDM_SEG_LEAN=$l timeout 120 python tools/seg_ab.py $k 128; done; SLDM_SEG_LEAN=$l timeout 120 python tools/seg_ab.py batch 64; done 2>&1 | grep -E "fwd_mean|Error" | tee gpurun_out/seg_ab_14.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_tc_paths_gpu.py tests/test_properties_gpu.py -x -q -m gpu 2>&1 | tail -4
