#!/bin/bash
# round 2: 8-GPU validation of the host seam (one process per GPU, 64 callers per GPU)
mkdir -p gpurun_out
nproc > gpurun_out/nproc8.txt; nvidia-smi -L | wc -l >> gpurun_out/nproc8.txt
export CSBWA_CO_TIMING=1
run() {  # name, nproc, extra env
  local name=$1 n=$2; shift 2
  t0=$(date +%s)
  if [ "$n" = "1" ]; then
    env "$@" timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-matesw --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  else
    env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  fi
  echo "$name rc=$? ($(( $(date +%s) - t0 )) s)"
  python -c "
import json;d=json.load(open('gpurun_out/bench_$name.json'))
print('$name value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),[round(x,1) for x in d['e2e']['repetitions_gcups']],'pageable',round(d['e2e']['pageable']['value'],1),'cpus',d['host_cpus'],'calls/sub',round(d['e2e_calls_per_device_submission'],2),'roofline frac',round(d['roofline']['frac'],3))"
  grep "csbwa coalescer\] calls" gpurun_out/bench_$name.err | tail -2
}
run 8gpu_sm 8 CSBWA_CO_COPY=sm
run 8gpu_dma 8 CSBWA_CO_COPY=dma
run 1gpu_on8 1 CSBWA_CO_COPY=sm
