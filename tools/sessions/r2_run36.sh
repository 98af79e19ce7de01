#!/bin/bash
# session 36: more callers than the bench's 64 (how close the seam gets to the resident figure when it has the calls)
cd /root/repo; mkdir -p gpurun_out
export PROBE_REPEAT=40
PROBE_CFGS="96 1 0 CSBWA_CO_SLOTS=24
128 1 0 CSBWA_CO_SLOTS=32
192 1 0 CSBWA_CO_SLOTS=32" bash tools/e2e_probe.sh > gpurun_out/s36_probe.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/s36_probe.log"):
    l=l.strip()
    if l.startswith("=="): print(l, end=" -> ")
    elif l.startswith("{"):
        d=json.loads(l); print("gcups", round(d["gcups"],1), "calls/group", round(d["calls_per_group"],2), "ms/group", d["ms_per_group"]["host_ms"], "ms/call", round(d["ms_per_call"],3))
PY
