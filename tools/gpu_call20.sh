#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/call20_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/call20_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/gpu_bench_all.sh r02g
