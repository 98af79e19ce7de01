#!/bin/bash
# round 2, GPU session 7: chain2aln pipeline + coords seam (parity + bench), ncu captures of the extension step
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m pytest tests/test_coords.py tests/test_chain2aln.py tests/test_jni_glue.py -m gpu -x -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
tail -3 gpurun_out/pytest_gpu7.log
CSBWA_C2A_TIMING=1 timeout 300 python tools/bench_chain2aln.py > gpurun_out/chain2aln.json 2> gpurun_out/chain2aln.err; echo "chain2aln rc=$?"; cat gpurun_out/chain2aln.json; tail -4 gpurun_out/chain2aln.err
timeout 300 python tools/bench_coords.py > gpurun_out/coords.json 2> gpurun_out/coords.err; echo "coords rc=$?"; cat gpurun_out/coords.json; tail -2 gpurun_out/coords.err
timeout 300 python tools/bench_coords.py --reads-per-call 4096 --threads 32 > gpurun_out/coords4096.json 2>> gpurun_out/coords.err; echo "coords4096 rc=$?"; cat gpurun_out/coords4096.json
# launch list of the resident step (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/r2_launches_ext.csv python bench.py --pairs 131072 --steps 2 --no-graph --streams 1 --no-cpu-baseline --no-e2e --no-matesw > /dev/null 2> gpurun_out/ncu1.err; echo "ncu launches rc=$?"
# the OVERLAPPED step as one workload: the whole CUDA graph of a step profiled as a unit
timeout 900 ncu --graph-profiling graph --clock-control none --metrics smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed -c 2 --csv --log-file gpurun_out/r2_ncu_graph_step.csv python bench.py --pairs 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-matesw > /dev/null 2> gpurun_out/ncu2.err; echo "ncu graph rc=$?"
tail -5 gpurun_out/r2_ncu_graph_step.csv | cut -c1-1200
tail -3 gpurun_out/ncu2.err
