#!/bin/bash
# round 2, GPU session 4: full GPU suite, one-warp-block classes experiment
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
tail -3 gpurun_out/pytest_gpu4.log
for k in 2 3 4 6; do
  for wl in C2 C5; do
  CSBWA_EXT_ONE_WARP_CLS=$k timeout 300 python bench.py --workload $wl --no-e2e --no-matesw --no-cpu-baseline --steps 10 > gpurun_out/bench_ow${k}_$wl.json 2> gpurun_out/bench_ow${k}_$wl.err; echo "bench $wl onewarp<=$k rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_ow${k}_$wl.json'));print('$wl', $k, round(d['value'],1), 'GCUPS frac', round(d['roofline']['frac'],3), d['roofline']['phase_ms_sample'])"
  done
done
