#!/bin/bash
# session 22: mate-SW seam, groups in flight for 10-pair calls
cd /root/repo; mkdir -p gpurun_out
for s in 6 12 24; do
  echo "== aln slots $s"
  CSBWA_ALN_CO_SLOTS=$s timeout 300 python tools/bench_matesw.py --configs C3 --pairs 32768 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('kernel', round(d.get('kernel_gcups',0),1), 'large', round(d['host_abi_large']['gcups'],1), 'sbatch10', round(d['host_abi_sbatch10']['gcups'],1), 'calls/sub', round(d['host_abi_sbatch10']['calls_per_device_submission'],1))"
done
