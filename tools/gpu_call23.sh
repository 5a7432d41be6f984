#!/bin/bash
# round 2, call 23: CSR build A/B after the look-back / histogram / rowptr changes + per-kernel launch list
set -uo pipefail
mkdir -p gpurun_out
out=gpurun_out/ab_csr_23.jsonl; : > $out
timeout 300 python tools/ab_csr.py batch c4 c1 mid >> $out 2>gpurun_out/ab_csr_23.err
SLDM_LIB_PATH=build/ab/os3.so timeout 300 python tools/ab_csr.py batch c4 c1 mid >> $out 2>>gpurun_out/ab_csr_23.err
cat $out | cut -c1-140
for db in 8 10; do
  SLDM_CSR_DIGIT_BITS=$db timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_readout.py tests/test_properties_gpu.py -x -q -m gpu -k "csr or readout or propert" 2>&1 | tail -2
done
for w in batch c4; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02i_csr_${w}.csv \
    python tools/prof_csr.py $w > gpurun_out/ncu_csr_$w.log 2>&1
  tail -1 gpurun_out/ncu_csr_$w.log
done
tail -5 gpurun_out/ab_csr_23.err
