#!/bin/bash
# tools/gpu_scale.sh N TAG : the default workload on N GPUs, fp32 and bf16 features (gpurun --gpus N)
set -uo pipefail
n=$1; tag=${2:-r02}
mkdir -p gpurun_out
for dt in f32 bf16; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 --dtype $dt --no-cpu \
    > gpurun_out/${tag}_bench_batch_${n}gpu_${dt}.json 2> gpurun_out/${tag}_bench_batch_${n}gpu_${dt}.err
  echo "N=$n $dt rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_batch_${n}gpu_${dt}.json')); print('  ms', round(d['ms_per_step'],3), 'value %.3e'%d['value'], 'graphs/s %.3e'%d['graphs_per_sec'], 'e2e ms', round(d['e2e']['ms_per_step'],3), 'e2e graphs/s %.3e'%d['e2e']['graphs_per_sec'], d.get('grad_sync_check'))"
done
