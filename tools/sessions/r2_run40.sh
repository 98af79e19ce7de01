#!/bin/bash
# round 2, GPU session 40: insertion chain kept packed (CSBWA_P2_VARIANT 1) / with the shifts on the dot-product unit
# (3) against the shipped pair step (0): A/B on resident inputs, then the extension parity tests on the winner
mkdir -p gpurun_out
P=cloud-scale-bwamem_b200
run() {   # workload variant
  lib=$PWD/$P/libcsbwa_sw.so; [ $2 != 0 ] && lib=$PWD/$P/libcsbwa_sw_v$2.so
  CSBWA_LIB_PATH=$lib timeout 100 python bench.py --workload $1 --pairs 262144 --steps 10 --no-e2e --no-matesw --no-cpu-baseline > gpurun_out/s40_$1_v$2.json 2> gpurun_out/s40_$1_v$2.err
  python -c "import json;print(round(json.load(open('gpurun_out/s40_$1_v$2.json'))['value'],1))" 2>/dev/null || echo 0
}
best=0; bestv=0
for v in 0 1 3; do
  g=$(run C2 $v); echo "C2 variant $v: $g GCUPS"
  if python -c "import sys; sys.exit(0 if float('$g') > float('$best') else 1)"; then best=$g; bestv=$v; fi
done
echo "best on C2: variant $bestv ($best)"
if [ $bestv != 0 ]; then
  echo "C1 variant 0: $(run C1 0) GCUPS"; echo "C1 variant $bestv: $(run C1 $bestv) GCUPS"
  t0=$(date +%s)
  CSBWA_LIB_PATH=$PWD/$P/libcsbwa_sw_v$bestv.so timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or random_and_adversarial or zdrop or workloads_full" > gpurun_out/s40_pytest_v$bestv.log 2>&1
  echo "pytest variant $bestv rc=$? ($(( $(date +%s) - t0 )) s)"; tail -2 gpurun_out/s40_pytest_v$bestv.log
fi
