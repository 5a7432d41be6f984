#!/bin/bash
# round 2, call 25: full GPU suite after the CSR rewrite, smoke, host-side profile of the c1 step, bench lines batch / infer / c1
set -uo pipefail
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -x -q -m gpu) > gpurun_out/call25_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/call25_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python tools/host_profile.py > gpurun_out/r02k_host_profile_c1.txt 2>&1; head -45 gpurun_out/r02k_host_profile_c1.txt
for w in batch infer c1; do
  timeout 400 python bench.py --workload $w > gpurun_out/r02k_bench_$w.json 2> gpurun_out/r02k_bench_$w.err; echo "$w rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r02k_bench_$w.json"))
print({k:d.get(k) for k in ("ms_per_step","value","gpu_launches","host_enqueue_ms_per_step")}, d["e2e"].get("ms_per_step"), {k:v.get("ms") for k,v in d.get("kernels",{}).items()})
PY
done
