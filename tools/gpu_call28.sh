#!/bin/bash
# round 2, call 28: programmatic dependent launch on the hot-path kernels: full suite, then A/B of the step with the attribute on / off
set -uo pipefail
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -x -q -m gpu) > gpurun_out/call28_tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed" gpurun_out/call28_tests.log | tail -2
for rep in 1 2; do for pdl in 0 1; do
  for w in batch c1 infer; do
    SLDM_DISABLE_PDL=$pdl timeout 300 python bench.py --workload $w --no-cpu --no-c4 --skip-kernel-timing 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); g=d.get('cuda_graph') or {}
print('disable_pdl=$pdl', '$w', 'ms', round(d['ms_per_step'],4), 'graph', (g.get('inference') or {}).get('ms_per_step'), (g.get('training') or {}).get('ms_per_step'))"
  done
done; done
