#!/bin/bash
# round 2, call 33: per-kernel times of the readout and of the map attention (forward + backward) at the bench shape
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02r_readout_launches.csv python tools/prof_readout.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02r_mapatt_launches.csv python tools/ma_prof.py > /dev/null 2>&1
python - <<'PY'
import csv
for n in ("readout","mapatt"):
    rows=[r for r in csv.reader(l for l in open(f"gpurun_out/r02r_{n}_launches.csv") if l.startswith('"'))]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
    ks=[(r[ki][:70], float(r[vi].replace(',',''))/1000) for r in rows[1:]]
    per=len(ks)//3
    print(n, len(ks), "last iteration:")
    for k,t in ks[-per:]: print("   %-72s %8.1f us"%(k,t))
    print("   total %.1f us"%sum(t for _,t in ks[-per:]))
PY
