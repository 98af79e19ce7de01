#!/bin/bash
# session 30: 2-D sort key of the both-sides pass (max/8, side, min/4)
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ext_ or large" > gpurun_out/s30_tests.log 2>&1; tail -2 gpurun_out/s30_tests.log
export PROBE_REPEAT=30 CSBWA_CO_TRACE=1
PROBE_CFGS="64 1 0
64 1 0
16 1 0" bash tools/e2e_probe.sh > gpurun_out/s30_probe.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/s30_probe.log"):
    l=l.strip()
    if l.startswith("=="): print(l)
    elif l.startswith("{"):
        d=json.loads(l); print("   gcups", round(d["gcups"],1), "calls/group", round(d["calls_per_group"],2), "ms/group", d["ms_per_group"]["host_ms"])
    elif "device phases" in l: print("   ", l[18:])
PY
unset CSBWA_CO_TRACE
for wl in C2 C1; do
  echo -n "$wl resident, both-sides pass forced: "
  CSBWA_EXT_FUSED_MAX=1000000000 python bench.py --pairs 500000 --steps 5 --warmup 3 --no-e2e --no-matesw --no-cpu-baseline --workload $wl 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['unit'])"
done
