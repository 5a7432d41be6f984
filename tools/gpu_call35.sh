#!/bin/bash
# round 2, call 35: k_sage_tc walking its row blocks backwards (SLDM_TC_REVERSE bit 0 forward projection, bit 1 dgrad): the rows the previous kernel wrote last are read first
mkdir -p gpurun_out
SLDM_TC_REVERSE=3 timeout 600 python -m pytest tests/test_tc_paths_gpu.py tests/test_gpu_parity.py tests/test_bench_shapes_gpu.py -x -q -m gpu 2>&1 | tail -2
for rep in 1 2 3; do for m in 0 1 2 3; do
  SLDM_TC_REVERSE=$m timeout 300 python bench.py --no-cpu --no-c4 --skip-kernel-timing 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('reverse=$m', 'ms', round(d['ms_per_step'],4))"
done; done
for m in 0 1; do SLDM_TC_REVERSE=$m timeout 300 python bench.py --workload infer --no-cpu --skip-kernel-timing 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('infer reverse=$m', 'ms', round(d['ms_per_step'],4))"; done
