#!/bin/bash
# round 2, GPU session 1: parity suite, host-seam probes, default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
tail -5 gpurun_out/pytest_gpu.log
export PROBE_REPEAT=60
PROBE_CFGS="64 1 0
64 1 4
64 0 0
64 0 4
16 1 0
128 1 0
64 1 0 CSBWA_CO_INFLIGHT=2
64 1 0 CSBWA_CO_INFLIGHT=4
64 1 0 CSBWA_CO_INFLIGHT=8
64 1 4 CSBWA_CO_INFLIGHT=4
64 1 0 CSBWA_CO_GRAPH=0" bash tools/e2e_probe.sh > gpurun_out/probe1.log 2>&1
cat gpurun_out/probe1.log
t0=$(date +%s)
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
cut -c1-1500 gpurun_out/bench.json
tail -5 gpurun_out/bench.err
