#!/bin/bash
# session 26: batched copies (one driver call per direction and group)
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "coalescer or concurrent or pinned or callback" > gpurun_out/s26_tests.log 2>&1
tail -3 gpurun_out/s26_tests.log
export PROBE_REPEAT=30 CSBWA_CO_TRACE=1
PROBE_CFGS="64 1 0 CSBWA_CO_BATCHCOPY=0
64 1 0
64 1 0 CSBWA_CO_BATCHCOPY=0
64 1 0
64 0 0 CSBWA_CO_BATCHCOPY=0
64 0 0
64 1 4 CSBWA_CO_BATCHCOPY=0
64 1 4
16 1 0" bash tools/e2e_probe.sh > gpurun_out/s26_probe.log 2>&1
python - <<'PY'
import json,re
cur=None
for l in open("gpurun_out/s26_probe.log"):
    l=l.strip()
    if l.startswith("=="): cur=l; print(cur)
    elif l.startswith("{"):
        d=json.loads(l); print("   gcups", round(d["gcups"],1), "calls/group", round(d["calls_per_group"],2), "ms/group", d["ms_per_group"]["host_ms"])
    elif "per group us" in l or "device phases" in l: print("   ", l[l.index("per group")-0:] if "per group us" in l else l[18:])
PY
