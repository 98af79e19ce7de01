#!/bin/bash
# session 33: preparation kernels at the highest launch priority
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ext_golden or adversarial or coalescer or device" > gpurun_out/s33_tests.log 2>&1; tail -2 gpurun_out/s33_tests.log
export PROBE_REPEAT=30 CSBWA_CO_TRACE=1
PROBE_CFGS="64 1 0 CSBWA_EXT_PREP_PRIO=0
64 1 0
64 1 0 CSBWA_EXT_PREP_PRIO=0
64 1 0
16 1 0 CSBWA_EXT_PREP_PRIO=0
16 1 0" bash tools/e2e_probe.sh > gpurun_out/s33_probe.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/s33_probe.log"):
    l=l.strip()
    if l.startswith("=="): print(l)
    elif l.startswith("{"):
        d=json.loads(l); print("   gcups", round(d["gcups"],1), "calls/group", round(d["calls_per_group"],2), "ms/group", d["ms_per_group"]["host_ms"])
    elif "device phases" in l: print("   ", l[18:])
PY
