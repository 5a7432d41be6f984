#!/bin/bash
# round 2, call 29: final bench lines (r02n) + ncu launch list of the default step + full-set capture of the CSR kernels
set -uo pipefail
mkdir -p gpurun_out
bash tools/gpu_bench_all.sh r02n
timeout 300 python bench.py --workload c2 --c2-graphs 64 --no-cpu > gpurun_out/r02n_bench_c2_64.json 2> gpurun_out/r02n_bench_c2_64.err; echo "c2_64 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02n_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-c4 --skip-kernel-timing > gpurun_out/ncu_launch.log 2>&1
tail -c 200 gpurun_out/ncu_launch.log; echo
ncu --set full --clock-control none --import-source on -k regex:"k_onesweep_pass|k_convert_hist|k_rowptr_from_sorted" -s 15 -c 5 \
    -o gpurun_out/r02n_csr_full -f python tools/prof_csr.py batch > gpurun_out/ncu_csr_full.log 2>&1
tail -2 gpurun_out/ncu_csr_full.log
