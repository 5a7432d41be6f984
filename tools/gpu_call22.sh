#!/bin/bash
# round 2, call 22: per-kernel times of the new CSR build (ncu launch list; cold-cache, serialised) + full-set capture of one pass
set -uo pipefail
mkdir -p gpurun_out
for w in batch c4; do
  SLDM_CSR_DIGIT_BITS=8 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02h_csr_${w}.csv \
    python tools/prof_csr.py $w > gpurun_out/ncu_csr_$w.log 2>&1
  tail -2 gpurun_out/ncu_csr_$w.log
done
SLDM_CSR_DIGIT_BITS=8 ncu --set full --clock-control none --import-source on -k regex:"k_onesweep_pass|k_convert_hist|k_rowptr" -s 18 -c 6 \
    -o gpurun_out/r02h_csr_full -f python tools/prof_csr.py batch > gpurun_out/ncu_csr_full.log 2>&1
tail -2 gpurun_out/ncu_csr_full.log
