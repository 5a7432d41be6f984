# profile pass: ncu launch list of the bench command (whole steps only), ncu --set full of the kernel groups
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r01i_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --skip-kernel-timing > gpurun_out/ncu_launch.log 2>&1
tail -c 300 gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:"k_segment_rows_lean|k_sage_tc|k_wgrad_tc|k_ln_bwd_rows|k_onesweep_pass|k_reduce_parts" -c 26 \
    -o gpurun_out/r01i_full -f python tools/prof_kernels.py all > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
