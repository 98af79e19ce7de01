#!/bin/bash
# round 2, GPU session 8: coords coalescing, ncu evidence, C1/C5 lines, C3 at 1M pairs
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m pytest tests/test_coords.py tests/test_chain2aln.py tests/test_jni_glue.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
tail -3 gpurun_out/pytest_gpu8.log
CSBWA_C2A_TIMING=1 timeout 300 python tools/bench_chain2aln.py > gpurun_out/chain2aln.json 2> gpurun_out/chain2aln.err; echo "chain2aln rc=$?"; cat gpurun_out/chain2aln.json; tail -2 gpurun_out/chain2aln.err
timeout 300 python tools/bench_coords.py > gpurun_out/coords.json 2> gpurun_out/coords.err; echo "coords rc=$?"; cat gpurun_out/coords.json; tail -2 gpurun_out/coords.err
timeout 300 python tools/bench_coords.py --reads-per-call 4096 --threads 32 > gpurun_out/coords4096.json 2>> gpurun_out/coords.err; echo "coords4096 rc=$?"; cat gpurun_out/coords4096.json
export CSBWA_PROFILE_STEP=1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_ext.csv python bench.py --pairs 131072 --steps 1 --no-graph --streams 1 --no-cpu-baseline --no-e2e --no-matesw > /dev/null 2> gpurun_out/ncu1.err; echo "ncu launches rc=$?"
timeout 900 ncu --profile-from-start off --graph-profiling graph --clock-control none --metrics smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -c 2 --csv --log-file gpurun_out/r2_ncu_graph_step.csv python bench.py --pairs 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-matesw > /dev/null 2> gpurun_out/ncu2.err; echo "ncu graph rc=$?"
grep -v "^==" gpurun_out/r2_ncu_graph_step.csv | cut -d, -f5,13- | head -24
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_ext_side -s 10 -c 4 -o gpurun_out/r2_ext_side_full python bench.py --pairs 131072 --steps 1 --no-graph --streams 1 --no-cpu-baseline --no-e2e --no-matesw > /dev/null 2> gpurun_out/ncu3.err; echo "ncu full rc=$?"
unset CSBWA_PROFILE_STEP
for wl in C1 C5; do
  timeout 400 python bench.py --workload $wl --no-matesw --no-cpu-baseline --steps 10 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_$wl.json'));print('$wl', round(d['value'],1), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'pageable', round(d['e2e']['pageable']['value'],1))"
done
timeout 600 python tools/bench_matesw.py --configs C3 --pairs 1000000 --no-host-legs > gpurun_out/matesw_C3_1M.json 2> gpurun_out/matesw_C3_1M.err; echo "C3 1M rc=$?"; cut -c1-900 gpurun_out/matesw_C3_1M.json; tail -2 gpurun_out/matesw_C3_1M.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; cut -c1-400 gpurun_out/bench_reference.json
