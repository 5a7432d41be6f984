#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/call10_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/call10_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/r02a_bench_batch.json 2> gpurun_out/r02a_bench_batch.err; echo "bench rc=$?"; head -c 400 gpurun_out/r02a_bench_batch.json; echo
timeout 300 python bench.py --dtype bf16 --no-cpu > gpurun_out/r02a_bench_batch_bf16.json 2> gpurun_out/r02a_bench_batch_bf16.err; echo "bench bf16 rc=$?"; head -c 400 gpurun_out/r02a_bench_batch_bf16.json; tail -3 gpurun_out/r02a_bench_batch_bf16.err; echo
timeout 300 python bench.py --workload infer --no-cpu > gpurun_out/r02a_bench_infer.json 2>/dev/null; head -c 300 gpurun_out/r02a_bench_infer.json; echo
timeout 300 python bench.py --workload infer --dtype bf16 --no-cpu > gpurun_out/r02a_bench_infer_bf16.json 2>/dev/null; head -c 300 gpurun_out/r02a_bench_infer_bf16.json; echo
timeout 300 python bench.py --workload c1 --no-cpu > gpurun_out/r02a_bench_c1.json 2>/dev/null; head -c 300 gpurun_out/r02a_bench_c1.json; echo
