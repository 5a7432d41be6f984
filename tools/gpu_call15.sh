#!/bin/bash
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/call15_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/call15_tests.log
for w in "c1" "c2 --c2-graphs 32"; do timeout 300 python bench.py --workload $w --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config'].get('workload'), 'ms', round(d['ms_per_step'],3), 'host', d.get('host_enqueue_ms_per_step'), 'launches', d.get('gpu_launches'))"; done
timeout 200 python tools/host_profile.py 2>&1 | tail -25
