#!/bin/bash
# session 15: narrow-band fast path of the lane-group row; busy policy of the seam
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ext_ or large" > gpurun_out/s15_tests.log 2>&1
tail -3 gpurun_out/s15_tests.log
export PROBE_REPEAT=20
PROBE_CFGS="1 1 0
1 1 0 CSBWA_EXT_COOP_G=16
2 1 0
2 1 0 CSBWA_EXT_COOP_BUSY=1
3 1 0
3 1 0 CSBWA_EXT_COOP_BUSY=1
4 1 0
4 1 0 CSBWA_EXT_COOP_BUSY=1
8 1 0
8 1 0 CSBWA_EXT_COOP_BUSY=1
64 1 0
64 1 0 CSBWA_EXT_COOP_BUSY=1" bash tools/e2e_probe.sh > gpurun_out/s15_probe.log 2>&1
grep -E "^==|gcups" gpurun_out/s15_probe.log | sed -E 's/.*"gcups": ([0-9.]+), "calls_per_group": ([0-9.]+), "ms_per_group": \{"host_ms": ([0-9.]+).*/  gcups \1 calls\/group \2 ms\/group \3/'
