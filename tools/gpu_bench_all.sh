#!/bin/bash
# every bench line of DESIGN.md section 5 in one call: tools/gpurun_retry.sh 1800 'bash tools/gpu_bench_all.sh r02d'
set -uo pipefail
tag=${1:-r02}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/${tag}_bench_${name}.json 2> gpurun_out/${tag}_bench_${name}.err; echo "$name rc=$? $(head -c 230 gpurun_out/${tag}_bench_${name}.json | cut -d, -f2,8)"; }
run batch
run batch_bf16 --dtype bf16 --no-cpu
run infer --workload infer --no-cpu
run infer_bf16 --workload infer --dtype bf16 --no-cpu
run c4 --workload c4 --no-cpu
run c1 --workload c1
run c2_32 --workload c2 --c2-graphs 32 --no-cpu
run c2_1024 --workload c2 --c2-graphs 1024 --no-cpu
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>/dev/null; head -c 200 gpurun_out/${tag}_bench_reference.json; echo
