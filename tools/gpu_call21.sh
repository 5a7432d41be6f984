#!/bin/bash
# round 2, call 21: CSR build A/B (previous head vs fused convert+hist / two-sort launches / wide look-back; 8- vs 10-bit digits; 3 vs 4 CTAs per SM)
set -uo pipefail
mkdir -p gpurun_out
out=gpurun_out/ab_csr_21.jsonl; : > $out
SLDM_LIB_PATH=build/ab/r02head.so timeout 300 python tools/ab_csr.py batch c4 c1 mid >> $out 2>gpurun_out/ab_csr_21.err
for db in 8 10; do
  SLDM_CSR_DIGIT_BITS=$db timeout 300 python tools/ab_csr.py batch c4 c1 mid >> $out 2>>gpurun_out/ab_csr_21.err
  SLDM_CSR_DIGIT_BITS=$db SLDM_LIB_PATH=build/ab/os3.so timeout 300 python tools/ab_csr.py batch c4 c1 mid >> $out 2>>gpurun_out/ab_csr_21.err
done
cat $out
for db in 8 10; do
  SLDM_CSR_DIGIT_BITS=$db timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_readout.py tests/test_properties_gpu.py -x -q -m gpu -k "csr or readout or propert" 2>&1 | tail -3
done
tail -5 gpurun_out/ab_csr_21.err
