#!/bin/bash
# session 23: compute-sanitizer over the build with the lane-group kernel (memcheck, racecheck, synccheck)
cd /root/repo; mkdir -p gpurun_out
S=gpurun_out/r2_sanitizer.txt
echo "# compute-sanitizer on a B200 (round 2 build with k_ext_small; smoke()'s extension call and the golden set take the lane-group kernel)" > $S
echo "## memcheck: python -c 'import __graft_entry__ as g; g.smoke()'" >> $S
timeout 900 compute-sanitizer --tool memcheck python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | grep -v "^=========     \|Warning" | tail -6 >> $S
echo "## memcheck: pytest tests/test_gpu_parity.py tests/test_global.py tests/test_coords.py tests/test_matesw_ref.py -m gpu -k 'adversarial or golden or pinned or callback or small_calls or global or coords_seam or native'" >> $S
timeout 1800 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_global.py tests/test_coords.py tests/test_matesw_ref.py -m gpu -q -k 'adversarial or golden or pinned or callback or small_calls or global or coords_seam or native' 2>&1 | grep -v "^=========     \|Warning" | tail -8 >> $S
echo "## racecheck: smoke()" >> $S
timeout 900 compute-sanitizer --tool racecheck python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | grep -v "^=========     \|Warning" | tail -5 >> $S
echo "## racecheck: pytest tests/test_gpu_parity.py -m gpu -k 'ext_golden' (class kernels, lane-group kernel, u8 core)" >> $S
timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k 'ext_golden' 2>&1 | grep -v "^=========     \|Warning" | tail -6 >> $S
echo "## synccheck: pytest tests/test_gpu_parity.py -m gpu -k 'ext_golden'" >> $S
timeout 900 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k 'ext_golden' 2>&1 | grep -v "^=========     \|Warning" | tail -6 >> $S
cat $S
