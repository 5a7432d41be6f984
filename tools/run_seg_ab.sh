mkdir -p gpurun_out
{
for C in 8 6 5 4 3 2; do
SLDM_SEG_CTAS=$C python tools/seg_ab.py batch 128 2>&1 | tail -1 | sed "s/^/ctas=$C /"
done
for C in 8 4; do SLDM_SEG_CTAS=$C python tools/seg_ab.py c4 128 2>&1 | tail -1 | sed "s/^/ctas=$C /"; done
for C in 8 4; do SLDM_SEG_CTAS=$C python tools/seg_ab.py batch 64 2>&1 | tail -1 | sed "s/^/ctas=$C /"; done
} > gpurun_out/seg_ab4.log 2>&1
cat gpurun_out/seg_ab4.log
