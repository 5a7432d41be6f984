#!/bin/bash
# round 2, call 24: CSR build: MATCH.ANY every 2nd / 3rd / 4th / never ranking round; src histogram split out again
set -uo pipefail
mkdir -p gpurun_out
out=gpurun_out/ab_csr_24.jsonl; : > $out
for lib in "" build/ab/me2.so build/ab/me4.so build/ab/me99.so; do
  SLDM_LIB_PATH=$lib timeout 300 python tools/ab_csr.py batch c4 c1 mid >> $out 2>>gpurun_out/ab_csr_24.err
done
cat $out | cut -c1-140
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_readout.py tests/test_properties_gpu.py -x -q -m gpu -k "csr or readout or propert" 2>&1 | tail -2
for w in batch c4; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02j_csr_${w}.csv \
    python tools/prof_csr.py $w > gpurun_out/ncu_csr_$w.log 2>&1
  tail -1 gpurun_out/ncu_csr_$w.log
done
tail -5 gpurun_out/ab_csr_24.err
