#!/bin/bash
# round 2, GPU session 10: SWGlobal nibble directions, compute-sanitizer, coordinate seam on a human-sized reference
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests/test_global.py -m gpu -x -q > gpurun_out/pytest_gpu10.log 2>&1; echo "pytest global rc=$? ($(( $(date +%s) - t0 )) s)"; tail -2 gpurun_out/pytest_gpu10.log
timeout 300 python tools/bench_global.py > gpurun_out/global.json 2> gpurun_out/global.err; echo "global rc=$?"; cut -c1-700 gpurun_out/global.json; tail -2 gpurun_out/global.err
timeout 600 ncu --set full --clock-control none -k regex:k_glb -s 1 -c 1 --csv --page raw --log-file gpurun_out/r2_ncu_glb_raw.csv python tools/bench_global.py --pairs 32768 --steps 1 > /dev/null 2> gpurun_out/ncu_glb.err; echo "ncu glb rc=$?"
S=gpurun_out/r2_sanitizer.txt
echo "# compute-sanitizer on a B200 (round 2 build)" > $S
echo "## memcheck: python -c 'import __graft_entry__ as g; g.smoke()'" >> $S
timeout 900 compute-sanitizer --tool memcheck python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | grep -v "^=========     \|Warning" | tail -6 >> $S
echo "## memcheck: pytest tests/test_gpu_parity.py tests/test_global.py tests/test_coords.py tests/test_matesw_ref.py -m gpu -k 'adversarial or golden or pinned or callback or small_calls or global or coords_seam or native'" >> $S
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_global.py tests/test_coords.py tests/test_matesw_ref.py -m gpu -q -k 'adversarial or golden or pinned or callback or small_calls or global or coords_seam or native' 2>&1 | grep -v "^=========     \|Warning" | tail -8 >> $S
echo "## racecheck: smoke()" >> $S
timeout 900 compute-sanitizer --tool racecheck python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | grep -v "^=========     \|Warning" | tail -5 >> $S
cat $S
timeout 900 python tools/bench_coords.py --ref-bp 3100000000 --pairs 131072 --reads-per-call 32768 > gpurun_out/coords_3g.json 2> gpurun_out/coords_3g.err; echo "coords 3.1G rc=$?"; cat gpurun_out/coords_3g.json; tail -2 gpurun_out/coords_3g.err
