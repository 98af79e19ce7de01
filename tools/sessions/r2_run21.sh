#!/bin/bash
# session 21: full GPU suite on the current build, default bench line, side benches
cd /root/repo; mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s21_pytest.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/s21_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/s21_bench.json 2> gpurun_out/s21_bench.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
python -c "
import json;d=json.load(open('gpurun_out/s21_bench.json'))
print('value',round(d['value'],1),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value'],1),[round(x,1) for x in d['e2e']['repetitions_gcups']],'pageable',round(d['e2e']['pageable']['value'],1),'cpu',round(d['cpu_baseline']['value'],2),'matesw',{k:(round(v,1) if isinstance(v,float) else v) for k,v in d.get('matesw',{}).items() if not isinstance(v,dict)})"
t0=$(date +%s)
timeout 600 python tools/bench_matesw.py --configs C1,C3 --pairs 32768 > gpurun_out/s21_matesw.jsonl 2> gpurun_out/s21_matesw.err; echo "matesw rc=$? ($(( $(date +%s) - t0 )) s)"
python -c "
import json
for l in open('gpurun_out/s21_matesw.jsonl'):
    d=json.loads(l); print(d.get('config',{}).get('workload','')[:30], 'kernel', round(d.get('kernel_gcups',0),1), 'large', round(d['host_abi_large']['gcups'],1), 'sbatch10', round(d['host_abi_sbatch10']['gcups'],1), 'ref', d.get('reference_cpu'))"
timeout 300 python tools/bench_chain2aln.py --steps 6 > gpurun_out/s21_chain2aln.json 2> gpurun_out/s21_chain2aln.err; echo "chain2aln rc=$?"; cat gpurun_out/s21_chain2aln.json
