#!/bin/bash
# session 12: lane-group side kernels -- parity, then latency / throughput of the host seam by caller count
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ext_ or coalescer or device or large" > gpurun_out/s12_tests.log 2>&1
tail -5 gpurun_out/s12_tests.log
export PROBE_REPEAT=20
PROBE_CFGS="1 1 0 CSBWA_EXT_COOP_MAX=0
1 1 0 CSBWA_EXT_COOP_MAX=8192
1 1 0 CSBWA_EXT_COOP_MAX=8192 CSBWA_EXT_COOP_G=16
1 1 0 CSBWA_EXT_COOP_MAX=8192 CSBWA_EXT_COOP_G=32
4 1 0 CSBWA_EXT_COOP_MAX=0
4 1 0 CSBWA_EXT_COOP_MAX=8192
4 1 0 CSBWA_EXT_COOP_MAX=16384
16 1 0 CSBWA_EXT_COOP_MAX=0
16 1 0 CSBWA_EXT_COOP_MAX=8192
16 1 0 CSBWA_EXT_COOP_MAX=16384
16 1 0 CSBWA_EXT_COOP_MAX=16384 CSBWA_EXT_COOP_G=16
64 1 0 CSBWA_EXT_COOP_MAX=0
64 1 0 CSBWA_EXT_COOP_MAX=8192
64 1 0 CSBWA_EXT_COOP_MAX=16384" bash tools/e2e_probe.sh > gpurun_out/s12_probe.log 2>&1
cat gpurun_out/s12_probe.log
