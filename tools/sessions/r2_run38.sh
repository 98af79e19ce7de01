#!/bin/bash
# session 38: pump nap between polls while groups are in flight
cd /root/repo; mkdir -p gpurun_out
export PROBE_REPEAT=30
PROBE_CFGS="64 1 0 CSBWA_CO_NAP_US=20
64 1 0 CSBWA_CO_NAP_US=10
64 1 0 CSBWA_CO_NAP_US=5
64 1 0 CSBWA_CO_NAP_US=40
64 0 0 CSBWA_CO_NAP_US=20
64 0 0 CSBWA_CO_NAP_US=5" bash tools/e2e_probe.sh > gpurun_out/s38_probe.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/s38_probe.log"):
    l=l.strip()
    if l.startswith("=="): print(l, end=" -> ")
    elif l.startswith("{"):
        d=json.loads(l); print("gcups", round(d["gcups"],1), "calls/group", round(d["calls_per_group"],2), "ms/group", d["ms_per_group"]["host_ms"])
PY
