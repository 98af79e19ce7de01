#!/bin/bash
# session 37: row-aligned target stream (uniform refill): parity, resident lines
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ext_ or large" > gpurun_out/s37_tests.log 2>&1; tail -2 gpurun_out/s37_tests.log
for wl in C2 C1 C5; do
  echo -n "$wl resident: "
  python bench.py --steps 8 --warmup 3 --no-e2e --no-matesw --no-cpu-baseline --workload $wl 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['unit'], 'frac', round(d['roofline']['frac'],3))"
done
