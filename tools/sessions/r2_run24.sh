#!/bin/bash
# session 24: where a group's time goes on the device under load (CSBWA_CO_TRACE)
cd /root/repo; mkdir -p gpurun_out
export PROBE_REPEAT=30 CSBWA_CO_TRACE=1
PROBE_CFGS="1 1 0 CSBWA_EXT_COOP_MAX=0
16 1 0
64 1 0
128 1 0 CSBWA_CO_SLOTS=32" bash tools/e2e_probe.sh > gpurun_out/s24_probe.log 2>&1
grep -E "^==|gcups|device phases|per group us" gpurun_out/s24_probe.log | sed -E 's/.*"gcups": ([0-9.]+), "calls_per_group": ([0-9.]+), "ms_per_group": \{"host_ms": ([0-9.]+).*/  gcups \1 calls\/group \2 ms\/group \3/' | cut -c1-300
