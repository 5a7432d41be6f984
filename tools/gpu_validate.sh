#!/bin/bash
# One-call validation of the whole repo on the GPU box:
#   tools/gpurun_retry.sh 900 'bash tools/gpu_validate.sh'
# full GPU test suite, smoke(), the default bench line; outputs under gpurun_out/validate_*.  Exits non-zero when the
# tests or smoke() fail.
set -uo pipefail
mkdir -p gpurun_out
rc=0
(time timeout 600 python -m pytest tests -x -q -m gpu) > gpurun_out/validate_tests.log 2>&1 || rc=1
tail -6 gpurun_out/validate_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/validate_smoke.log 2>&1 || rc=1
tail -1 gpurun_out/validate_smoke.log
timeout 300 python bench.py > gpurun_out/validate_bench_batch.json 2> gpurun_out/validate_bench_batch.err || rc=1
head -c 300 gpurun_out/validate_bench_batch.json; echo
exit $rc
