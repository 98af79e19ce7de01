#!/bin/bash
# round 2, GPU session 3: masked-edge p2 core (parity + C1/C2/C5 resident), pageable staging experiments, box topology
mkdir -p gpurun_out
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" > gpurun_out/topo.txt 2>&1
nvidia-smi topo -m >> gpurun_out/topo.txt 2>&1
t0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_coords.py tests/test_chain2aln.py tests/test_matesw_ref.py tests/test_matesw_group.py -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"
tail -3 gpurun_out/pytest_gpu3.log
for wl in C2 C1 C5; do
  timeout 300 python bench.py --workload $wl --no-e2e --no-matesw --no-cpu-baseline --steps 10 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/bench_$wl.json'));print('$wl', round(d['value'],1), 'GCUPS frac', round(d['roofline']['frac'],3), d['roofline']['phase_ms_sample'])"
done
export PROBE_REPEAT=40
PROBE_CFGS="64 0 0
64 0 0 CSBWA_CO_NTCOPY=1
64 0 8 CSBWA_CO_NTCOPY=1
64 0 4 CSBWA_CO_NTCOPY=1
64 0 8
64 1 0" bash tools/e2e_probe.sh > gpurun_out/probe3.log 2>&1
cat gpurun_out/probe3.log | grep -v "graphs built"
cat gpurun_out/topo.txt
