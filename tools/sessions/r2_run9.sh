#!/bin/bash
# round 2, GPU session 9: hoisted row operands + refilling kernel with private chunks (parity + A/B)
mkdir -p gpurun_out
t0=$(date +%s)
CSBWA_EXT_REFILL=126 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ext or multi_call or device_resident or large" > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest(refill) rc=$? ($(( $(date +%s) - t0 )) s)"
tail -3 gpurun_out/pytest_gpu9.log
for rf in 0 126 6; do
  for wl in C2 C5 C1; do
  CSBWA_EXT_REFILL=$rf timeout 300 python bench.py --workload $wl --no-e2e --no-matesw --no-cpu-baseline --steps 10 > gpurun_out/bench_rf${rf}_$wl.json 2> gpurun_out/bench_rf${rf}_$wl.err; echo "bench $wl refill=$rf rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_rf${rf}_$wl.json'));print('$wl', $rf, round(d['value'],1), 'GCUPS frac', round(d['roofline']['frac'],3), d['roofline']['phase_ms_sample'])"
  done
done
