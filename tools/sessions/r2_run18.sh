#!/bin/bash
# session 18: 8 GPUs, one process per GPU, 64 callers per GPU, pump thread at nice -10
mkdir -p gpurun_out
nproc > gpurun_out/nproc8.txt; nvidia-smi -L | wc -l >> gpurun_out/nproc8.txt
export CSBWA_CO_TIMING=1
t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu_nice.json 2> gpurun_out/bench_8gpu_nice.err
echo "rc=$? ($(( $(date +%s) - t0 )) s)"
python -c "
import json;d=json.load(open('gpurun_out/bench_8gpu_nice.json'))
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),[round(x,1) for x in d['e2e']['repetitions_gcups']],'pageable',round(d['e2e']['pageable']['value'],1),'cpus',d['host_cpus'],'calls/sub',round(d['e2e_calls_per_device_submission'],2),'roofline frac',round(d['roofline']['frac'],3))"
grep "csbwa coalescer\] calls" gpurun_out/bench_8gpu_nice.err | tail -3
