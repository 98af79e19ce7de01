mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout 200 python tools/bench_global.py 2>/dev/null | tail -1 | tee gpurun_out/glb_ring_$i.json | cut -c1-330
done
CSBWA_GLB_NO_RING=1 timeout 200 python tools/bench_global.py 2>/dev/null | tail -1 | tee gpurun_out/glb_noring_1.json | cut -c1-330
