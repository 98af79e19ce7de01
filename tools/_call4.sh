mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_glb -s 1 -c 2 -f -o gpurun_out/glb2 python tools/bench_global.py --pairs 32768 --steps 1 > gpurun_out/ncu_glb2.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/ncu_glb2.log | cut -c1-300
ncu -i gpurun_out/glb2.ncu-rep --page raw --csv > gpurun_out/glb2_raw.csv 2>/dev/null
ncu -i gpurun_out/glb2.ncu-rep --page source --csv --kernel-name k_glb > gpurun_out/glb2_src.csv 2>/dev/null; wc -c gpurun_out/glb2_*.csv
