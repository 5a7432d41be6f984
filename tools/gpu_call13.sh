#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
CMD1="python bench.py --steps 2 --warmup 3 --no-cpu --skip-kernel-timing --no-c4"
$CMD1 > gpurun_out/ncu_plain1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $CMD1 > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
CMD2="python tools/prof_kernels.py all"
$CMD2 > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_sage_tc|k_wgrad_tc|k_ln_bwd_rows|k_segment_rows_lean' -c 12 -o gpurun_out/r02_prof_layer $CMD2 > gpurun_out/ncu2.log 2>&1
echo "full set rc=$?"; ls -la gpurun_out/*.ncu-rep 2>/dev/null; tail -3 gpurun_out/ncu2.log
