#!/bin/bash
# round 2, GPU session 41: ALU-pipe relief in the other pair steps -- extension variant 4 (pair counters and the
# last-zero key by IMAD), mate-SW systolic step variant 3, SWGlobal pair step variant 3 (one experimental build with all
# three) against the shipped build: A/B on resident inputs, then the parity tests on the experimental build
mkdir -p gpurun_out
P=cloud-scale-bwamem_b200
A=$PWD/$P/libcsbwa_sw.so; B=$PWD/$P/libcsbwa_sw_exp.so
for wl in C2 C1; do for v in A B; do
  lib=$A; [ $v = B ] && lib=$B
  CSBWA_LIB_PATH=$lib timeout 60 python bench.py --workload $wl --pairs 262144 --steps 10 --no-e2e --no-matesw --no-cpu-baseline > gpurun_out/s41_ext_${wl}_$v.json 2> gpurun_out/s41_ext_${wl}_$v.err
  echo "ext $wl $v: $(python -c "import json;print(round(json.load(open('gpurun_out/s41_ext_${wl}_$v.json'))['value'],1))" 2>/dev/null) GCUPS"
done; done
for v in A B; do
  lib=$A; [ $v = B ] && lib=$B
  CSBWA_LIB_PATH=$lib timeout 60 python tools/bench_matesw.py --configs C3,C1 --pairs 8192 --no-host-legs --cpu-sample-jobs 64 > gpurun_out/s41_matesw_$v.jsonl 2> gpurun_out/s41_matesw_$v.err
  python -c "
import json
for l in open('gpurun_out/s41_matesw_$v.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print('matesw', d['workload'][:2], '$v', round(d['kernel_gcups'],1), 'GCUPS parity', d.get('parity_sample_ok'))"
  CSBWA_LIB_PATH=$lib timeout 60 python tools/bench_global.py --pairs 65536 --cpu-sample-jobs 512 > gpurun_out/s41_global_$v.json 2> gpurun_out/s41_global_$v.err
  python -c "
import json
d=json.load(open('gpurun_out/s41_global_$v.json')); print('global $v', round(d['kernel_gcups'],1), 'GCUPS parity', d.get('parity_sample_ok'))"
done
t0=$(date +%s)
CSBWA_LIB_PATH=$B timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_global.py tests/test_matesw_group.py -m gpu -x -q -k "golden or random_and_adversarial or zdrop or workloads_full or aln or gpu_vs_oracle or ring or read_shaped or product_vs_oracle" > gpurun_out/s41_pytest_exp.log 2>&1
echo "pytest exp rc=$? ($(( $(date +%s) - t0 )) s)"; tail -2 gpurun_out/s41_pytest_exp.log
