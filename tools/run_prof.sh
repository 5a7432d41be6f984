# round-1 profile pass: plain bench, ncu launch list of the same command, ncu --set full of the kernel groups
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_batch.json 2> gpurun_out/bench_batch.err || exit 1
tail -c 600 gpurun_out/bench_batch.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01b_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_segment_rows|k_sage_tc|k_wgrad_tc|k_ln_bwd_rows" -c 12 \
    -o gpurun_out/r01b_full -f python tools/prof_kernels.py all > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
