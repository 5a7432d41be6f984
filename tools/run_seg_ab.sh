mkdir -p gpurun_out
{
for F in 128 96 64 32; do python tools/seg_ab.py batch $F 2>&1 | tail -1; done
python tools/seg_ab.py c4 128 2>&1 | tail -1
} > gpurun_out/seg_ab7.log 2>&1
cat gpurun_out/seg_ab7.log
