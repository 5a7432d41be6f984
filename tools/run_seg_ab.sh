mkdir -p gpurun_out
{
for P in 0 20 24 30 36 48; do
SLDM_SEG_PAD_KB=$P python tools/seg_ab.py batch 128 2>&1 | tail -1 | sed "s/^/pad=$P /"
done
} > gpurun_out/seg_ab5.log 2>&1
cat gpurun_out/seg_ab5.log
