#!/bin/bash
# session 25: both sides of a task in one thread (one phase per launch sequence): parity, seam latency, resident throughput
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ext_ or coalescer or device or large" > gpurun_out/s25_tests.log 2>&1
tail -3 gpurun_out/s25_tests.log
export PROBE_REPEAT=30 CSBWA_CO_TRACE=1
PROBE_CFGS="1 1 0 CSBWA_EXT_COOP_MAX=0 CSBWA_EXT_FUSED_MAX=0
1 1 0 CSBWA_EXT_COOP_MAX=0
16 1 0 CSBWA_EXT_FUSED_MAX=0
16 1 0
64 1 0 CSBWA_EXT_FUSED_MAX=0
64 1 0
64 0 0 CSBWA_EXT_FUSED_MAX=0
64 0 0
128 1 0 CSBWA_CO_SLOTS=32" bash tools/e2e_probe.sh > gpurun_out/s25_probe.log 2>&1
grep -E "^==|gcups|device phases" gpurun_out/s25_probe.log | sed -E 's/.*"gcups": ([0-9.]+), "calls_per_group": ([0-9.]+), "ms_per_group": \{"host_ms": ([0-9.]+).*/  gcups \1 calls\/group \2 ms\/group \3/' | cut -c1-200
unset CSBWA_CO_TRACE
B="python bench.py --pairs 500000 --steps 5 --warmup 3 --no-e2e --no-matesw --no-cpu-baseline"
for wl in C2 C5 C1; do
  for f in 0 1000000000; do
    echo -n "$wl resident, FUSED_MAX=$f: "
    CSBWA_EXT_FUSED_MAX=$f $B --workload $wl 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['unit'], 'frac', round(d['roofline']['frac'],3))"
  done
done
