#!/bin/bash
# session 32: group buffers / groups in flight at the new operating point
cd /root/repo; mkdir -p gpurun_out
export PROBE_REPEAT=30
PROBE_CFGS="64 1 0 CSBWA_CO_SLOTS=8
64 1 0 CSBWA_CO_SLOTS=12
64 1 0
64 1 0 CSBWA_CO_SLOTS=24
64 1 0 CSBWA_CO_SLOTS=32
64 1 0 CSBWA_CO_SLOTS=32 CSBWA_CO_INFLIGHT=20" bash tools/e2e_probe.sh > gpurun_out/s32_probe.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/s32_probe.log"):
    l=l.strip()
    if l.startswith("=="): print(l, end=" -> ")
    elif l.startswith("{"):
        d=json.loads(l); print("gcups", round(d["gcups"],1), "calls/group", round(d["calls_per_group"],2), "ms/group", d["ms_per_group"]["host_ms"])
PY
