#!/bin/bash
# One-call validation of the whole repo on the GPU box (about 2.5 GPU-minutes):
#   tools/gpurun_retry.sh 400 'bash tools/gpu_validate.sh'
# full GPU test suite, smoke(), the default bench line and the C2 line; outputs under gpurun_out/validate_*.
mkdir -p gpurun_out
(time timeout 150 python -m pytest tests -x -q -m gpu) > gpurun_out/validate_tests.log 2>&1
tail -4 gpurun_out/validate_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/validate_smoke.log
timeout 150 python bench.py > gpurun_out/validate_bench_batch.json 2> gpurun_out/validate_bench_batch.err
head -c 260 gpurun_out/validate_bench_batch.json; echo
timeout 150 python bench.py --workload c2 --no-cpu > gpurun_out/validate_bench_c2.json 2> gpurun_out/validate_bench_c2.err
head -c 260 gpurun_out/validate_bench_c2.json; echo
