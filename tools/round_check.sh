#!/bin/bash
# Round-end check on the GPU box: GPU parity suite, smoke, the default bench line (run through gpurun).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
tail -3 gpurun_out/pytest_gpu.log
t0=$(date +%s)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? ($(( $(date +%s) - t0 )) s)"
tail -1 gpurun_out/smoke.log
t0=$(date +%s)
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
cut -c1-400 gpurun_out/bench.json
