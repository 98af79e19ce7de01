#!/bin/bash
# session 28: one-block preparation of mid-size launch sequences
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_coords.py tests/test_chain2aln.py -x -q -m gpu > gpurun_out/s28_tests.log 2>&1
tail -3 gpurun_out/s28_tests.log
export PROBE_REPEAT=30 CSBWA_CO_TRACE=1
PROBE_CFGS="64 1 0 CSBWA_EXT_PREP1=0
64 1 0
64 1 0 CSBWA_EXT_PREP1=0
64 1 0
16 1 0 CSBWA_EXT_PREP1=0
16 1 0
1 1 0 CSBWA_EXT_COOP_MAX=0 CSBWA_EXT_PREP1=0
1 1 0 CSBWA_EXT_COOP_MAX=0" bash tools/e2e_probe.sh > gpurun_out/s28_probe.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/s28_probe.log"):
    l=l.strip()
    if l.startswith("=="): print(l)
    elif l.startswith("{"):
        d=json.loads(l); print("   gcups", round(d["gcups"],1), "calls/group", round(d["calls_per_group"],2), "ms/group", d["ms_per_group"]["host_ms"])
    elif "device phases" in l: print("   ", l[18:])
PY
