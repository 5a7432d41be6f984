#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
timeout 120 python tools/ab_tc.py batch default 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/ab_tc_12.jsonl
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/call12_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/call12_tests.log
timeout 300 python bench.py --workload c1 --no-cpu > gpurun_out/r02c_bench_c1.json 2> gpurun_out/r02c_bench_c1.err; python -c "
import json; d=json.load(open('gpurun_out/r02c_bench_c1.json')); print('c1 ms', d['ms_per_step'], 'host', d['host_enqueue_ms_per_step'], 'launches', d['gpu_launches']); print(d.get('cuda_graph'))"
