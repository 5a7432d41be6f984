#!/bin/bash
# round 2, call 31: ncu --set full of the final layer kernels at the batch shape (refreshes profiles/r02_ncu_full_summary.txt, taken before the A 3 / B 3 rings)
set -uo pipefail
mkdir -p gpurun_out
python tools/prof_kernels.py all > gpurun_out/ncu_plain31.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_sage_tc|k_wgrad_tc|k_ln_bwd_rows|k_segment_rows_lean' -c 12 -o gpurun_out/r02q_prof_layer -f python tools/prof_kernels.py all > gpurun_out/ncu31.log 2>&1
echo "full set rc=$?"; tail -2 gpurun_out/ncu31.log; cat gpurun_out/ncu_plain31.log
