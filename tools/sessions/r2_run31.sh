#!/bin/bash
# session 31: 13-bit both-sides sort key with the cheaper preparation; both-sides pass against the two passes on resident input
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ext_ or large or coalescer" > gpurun_out/s31_tests.log 2>&1; tail -2 gpurun_out/s31_tests.log
export PROBE_REPEAT=30 CSBWA_CO_TRACE=1
PROBE_CFGS="64 1 0
64 1 0
16 1 0" bash tools/e2e_probe.sh > gpurun_out/s31_probe.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/s31_probe.log"):
    l=l.strip()
    if l.startswith("=="): print(l)
    elif l.startswith("{"):
        d=json.loads(l); print("   gcups", round(d["gcups"],1), "calls/group", round(d["calls_per_group"],2), "ms/group", d["ms_per_group"]["host_ms"])
    elif "device phases" in l: print("   ", l[18:])
PY
unset CSBWA_CO_TRACE
for wl in C2 C1 C5; do
  for f in 65536 1000000000; do
    echo -n "$wl resident 1M pairs, FUSED_MAX=$f: "
    CSBWA_EXT_FUSED_MAX=$f python bench.py --steps 5 --warmup 3 --no-e2e --no-matesw --no-cpu-baseline --workload $wl 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['unit'], 'frac', round(d['roofline']['frac'],3))"
  done
done
