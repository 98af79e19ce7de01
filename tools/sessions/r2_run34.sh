#!/bin/bash
# session 34: N GPUs of one box (N from the first argument), one process per GPU, 64 callers per GPU
N=${1:-2}
mkdir -p gpurun_out
nproc > gpurun_out/nproc_$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_${N}gpu.json'))
print('N=$N value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),[round(x,1) for x in d['e2e']['repetitions_gcups']],'pageable',round(d['e2e']['pageable']['value'],1),'cpus',d['host_cpus'])"
