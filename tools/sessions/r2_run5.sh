#!/bin/bash
# round 2, GPU session 5: refilling side kernel (parity + A/B on C2/C1/C5)
mkdir -p gpurun_out
t0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_coords.py tests/test_chain2aln.py -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
tail -3 gpurun_out/pytest_gpu5.log
for rf in 1 0; do
  for wl in C2 C1 C5; do
  CSBWA_EXT_REFILL=$rf timeout 300 python bench.py --workload $wl --no-e2e --no-matesw --no-cpu-baseline --steps 10 > gpurun_out/bench_rf${rf}_$wl.json 2> gpurun_out/bench_rf${rf}_$wl.err; echo "bench $wl refill=$rf rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_rf${rf}_$wl.json'));print('$wl', $rf, round(d['value'],1), 'GCUPS frac', round(d['roofline']['frac'],3), d['roofline']['phase_ms_sample'])"
  done
done
