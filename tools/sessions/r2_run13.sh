#!/bin/bash
# session 13: fused small-group kernel -- parity, then latency of the host seam by caller count / lanes per task
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ext_ or coalescer or device or large" > gpurun_out/s13_tests.log 2>&1
tail -5 gpurun_out/s13_tests.log
export PROBE_REPEAT=20
PROBE_CFGS="1 1 0 CSBWA_EXT_COOP_MAX=0
1 1 0 CSBWA_EXT_COOP_G=8
1 1 0 CSBWA_EXT_COOP_G=16
1 1 0 CSBWA_EXT_COOP_G=32
2 1 0 CSBWA_EXT_COOP_MAX=0
2 1 0 CSBWA_EXT_COOP_G=8
2 1 0 CSBWA_EXT_COOP_G=16
2 1 0 CSBWA_EXT_COOP_G=32
4 1 0 CSBWA_EXT_COOP_MAX=0
4 1 0 CSBWA_EXT_COOP_G=8
4 1 0 CSBWA_EXT_COOP_G=16
4 1 0 CSBWA_EXT_COOP_G=16 CSBWA_EXT_COOP_BUSY=1
4 1 0 CSBWA_EXT_COOP_G=32
8 1 0 CSBWA_EXT_COOP_MAX=0
8 1 0 CSBWA_EXT_COOP_G=16
8 1 0 CSBWA_EXT_COOP_G=16 CSBWA_EXT_COOP_BUSY=1
16 1 0 CSBWA_EXT_COOP_G=16
64 1 0 CSBWA_EXT_COOP_G=16
64 1 0 CSBWA_EXT_COOP_MAX=0" bash tools/e2e_probe.sh > gpurun_out/s13_probe.log 2>&1
grep -E "^==|gcups" gpurun_out/s13_probe.log | sed -E 's/.*"gcups": ([0-9.]+), "calls_per_group": ([0-9.]+), "ms_per_group": \{"host_ms": ([0-9.]+).*/  gcups \1 calls\/group \2 ms\/group \3/'
