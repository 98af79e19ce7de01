#!/bin/bash
# session 14: fused small-group kernel, parity again + the latency sweep's key points
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ext_ or coalescer or device or large" > gpurun_out/s14_tests.log 2>&1
tail -5 gpurun_out/s14_tests.log
export PROBE_REPEAT=20
PROBE_CFGS="1 1 0 CSBWA_EXT_COOP_MAX=0
1 1 0 CSBWA_EXT_COOP_G=16
1 1 0 CSBWA_EXT_COOP_G=32
2 1 0 CSBWA_EXT_COOP_MAX=0
2 1 0 CSBWA_EXT_COOP_G=32
4 1 0 CSBWA_EXT_COOP_G=32
1 0 0 CSBWA_EXT_COOP_MAX=0
1 0 0 CSBWA_EXT_COOP_G=32" bash tools/e2e_probe.sh > gpurun_out/s14_probe.log 2>&1
grep -E "^==|gcups" gpurun_out/s14_probe.log | sed -E 's/.*"gcups": ([0-9.]+), "calls_per_group": ([0-9.]+), "ms_per_group": \{"host_ms": ([0-9.]+).*/  gcups \1 calls\/group \2 ms\/group \3/'
