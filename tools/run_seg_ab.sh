mkdir -p gpurun_out
{
for F in 128 96 64 32; do for L in 0 4 8; do
SLDM_SEG_LEAN=$L python tools/seg_ab.py batch $F 2>&1 | tail -1
done; done
for L in 0 4 8; do SLDM_SEG_LEAN=$L python tools/seg_ab.py c4 128 2>&1 | tail -1; done
} > gpurun_out/seg_ab3.log 2>&1
cat gpurun_out/seg_ab3.log
