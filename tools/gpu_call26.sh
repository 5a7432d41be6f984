#!/bin/bash
# round 2, call 26: GraphedGruSage tests, c2 bench at 32 / 64 graphs with the graphed step, e2e_cabi_host record
set -uo pipefail
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_grusage.py -x -q -m gpu 2>&1 | tail -25
for g in 32 64; do
  timeout 400 python bench.py --workload c2 --c2-graphs $g --no-cpu > gpurun_out/r02l_bench_c2_$g.json 2> gpurun_out/r02l_bench_c2_$g.err; echo "c2 $g rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r02l_bench_c2_$g.json"))
print({k:d.get(k) for k in ("ms_per_step","value","gpu_launches")}, d.get("cuda_graph"))
PY
  tail -3 gpurun_out/r02l_bench_c2_$g.err
done
timeout 400 python bench.py --workload c1 --no-cpu > gpurun_out/r02l_bench_c1.json 2> gpurun_out/r02l_bench_c1.err; echo "c1 rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r02l_bench_c1.json"))
print({k:d.get(k) for k in ("ms_per_step","value","gpu_launches")}, d.get("e2e_cabi_host"), d.get("cuda_graph"))
PY
timeout 400 python bench.py --no-cpu --no-c4 > gpurun_out/r02l_bench_batch.json 2> gpurun_out/r02l_bench_batch.err; echo "batch rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r02l_bench_batch.json"))
print({k:d.get(k) for k in ("ms_per_step","value","gpu_launches")}, d["e2e"]["ms_per_step"], d.get("e2e_cabi_host"))
PY
