#!/bin/bash
# session 27: full GPU suite, default bench line, C1 / C5 lines on the build with fused sides + batched copies
cd /root/repo; mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s27_pytest.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - t0 )) s)"; tail -3 gpurun_out/s27_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/s27_bench.json 2> gpurun_out/s27_bench.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
show() { python -c "
import json,sys;d=json.load(open(sys.argv[1]))
print(sys.argv[1],'value',round(d['value'],1),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value'],1),[round(x,1) for x in d['e2e']['repetitions_gcups']],'pageable',round(d['e2e']['pageable']['value'],1),'calls/sub',round(d['e2e_calls_per_device_submission'],2),'launch->done ms',round(d['e2e_ms_launch_to_done_per_submission'],3))" $1; }
show gpurun_out/s27_bench.json
for wl in C1 C5; do
  timeout 600 python bench.py --workload $wl --no-matesw --no-cpu-baseline --steps 10 > gpurun_out/s27_bench_$wl.json 2> gpurun_out/s27_bench_$wl.err; echo "bench $wl rc=$?"
  show gpurun_out/s27_bench_$wl.json
done
