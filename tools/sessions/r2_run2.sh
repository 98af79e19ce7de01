#!/bin/bash
# round 2, GPU session 2: copy-engine vs SM gather/scatter for the coalesced seam
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "knobs or pinned or callback or concurrent or errors" > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
tail -5 gpurun_out/pytest_gpu2.log
export PROBE_REPEAT=60
PROBE_CFGS="64 1 0 CSBWA_CO_COPY=dma
64 1 4 CSBWA_CO_COPY=dma
64 0 0 CSBWA_CO_COPY=dma
64 0 4 CSBWA_CO_COPY=dma
64 0 0 CSBWA_CO_COPY=sm
64 1 0 CSBWA_CO_COPY=dma CSBWA_CO_SLOTS=32
128 1 0 CSBWA_CO_COPY=dma CSBWA_CO_SLOTS=32
32 1 0 CSBWA_CO_COPY=dma
16 1 0 CSBWA_CO_COPY=dma
64 1 0 CSBWA_CO_COPY=dma CSBWA_CO_INFLIGHT=8
64 0 0 CSBWA_CO_COPY=dma CSBWA_CO_GRAPH=0" bash tools/e2e_probe.sh > gpurun_out/probe2.log 2>&1
cat gpurun_out/probe2.log
