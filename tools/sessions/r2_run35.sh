#!/bin/bash
# session 35: launch list of the host seam (graph nodes profiled one by one): which kernels a small group and a
# mid-size group run, with their device times (cold, serialised -- shares, not benchmark values)
cd /root/repo; mkdir -p gpurun_out
timeout 600 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none -s 60 -c 240 --csv --log-file gpurun_out/r2_launches_seam.csv \
  python tools/e2e_probe.py --pairs 60000 --threads 6 --pinned 1 --repeat 8 > gpurun_out/s35.log 2>&1
echo "rc=$?"; grep -c "k_ext" gpurun_out/r2_launches_seam.csv
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2_launches_seam.csv")) if len(r)>5 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split("(")[0][:60]; t=float(r[-1].replace(",",""))
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=t
for k,v in agg.items(): print("%4d launches %10.1f us total  %s"%(v[0], v[1]/1e3 if rows and rows[0][-2]=="ns" else v[1], k))
PY
