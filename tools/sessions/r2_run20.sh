#!/bin/bash
# session 20: chain -> alignment driver, number of read ranges in flight
cd /root/repo; mkdir -p gpurun_out
for k in 4 8 16; do
  echo "== ranges $k"
  CSBWA_C2A_RANGES=$k CSBWA_C2A_TIMING=1 timeout 300 python tools/bench_chain2aln.py --steps 6 --cpu-reads 1024 2> gpurun_out/s20_c2a_$k.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['reads_per_s']/1e6,2),'M reads/s', round(d['ms_per_batch'],2),'ms', d['parity_sample_ok'])"
  grep chain2aln gpurun_out/s20_c2a_$k.err | tail -2
done
