#!/bin/bash
# round 2, call 34: readout with tie counts from the forward: tests, launch list, bench kernel groups
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_readout.py tests/test_grusage.py tests/test_properties_gpu.py -x -q -m gpu 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02s_readout_launches.csv python tools/prof_readout.py > /dev/null 2>&1; SLDM_READOUT_TWO_KERNELS=1 timeout 300 python -m pytest tests/test_readout.py -x -q -m gpu 2>&1 | tail -1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/r02s_readout_launches.csv") if l.startswith('"'))]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
ks=[(r[ki][:70], float(r[vi].replace(',',''))/1000) for r in rows[1:]]
per=len(ks)//3
for k,t in ks[-per:]: print("   %-72s %8.1f us"%(k,t))
print("   total %.1f us"%sum(t for _,t in ks[-per:]))
PY
timeout 300 python bench.py --no-cpu --no-c4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print(d['ms_per_step'], {n:(k[n]['ms'],k[n]['frac_hbm']) for n in k if 'readout' in n or 'map' in n})"
