#!/bin/bash
# round 2, GPU session 6: mate-SW coalescing (parity + bench), full default bench
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
tail -3 gpurun_out/pytest_gpu6.log
timeout 600 python tools/bench_matesw.py --configs C1,C3 --pairs 32768 > gpurun_out/matesw.jsonl 2> gpurun_out/matesw.err; echo "matesw rc=$?"
cat gpurun_out/matesw.jsonl | cut -c1-1500; tail -3 gpurun_out/matesw.err
t0=$(date +%s)
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
python -c "
import json;d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['repetitions_gcups'],'pageable',d['e2e']['pageable']['value'])
print('matesw',d.get('matesw')); print('swglobal',d.get('swglobal')); print('cpu',d.get('cpu_baseline'))"
tail -3 gpurun_out/bench.err
