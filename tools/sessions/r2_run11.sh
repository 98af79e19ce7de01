#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_coords.py -m gpu -x -q > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu11.log
for wl in C2 C1 C5; do
  timeout 300 python bench.py --workload $wl --no-e2e --no-matesw --no-cpu-baseline --steps 10 > gpurun_out/bench_pf_$wl.json 2> gpurun_out/bench_pf_$wl.err; echo "bench $wl rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_pf_$wl.json'));print('$wl', round(d['value'],1), 'GCUPS frac', round(d['roofline']['frac'],3), d['roofline']['phase_ms_sample'])"
done
