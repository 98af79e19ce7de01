#!/bin/bash
# session 19: what the 8-GPU box gives: topology, NUMA, aggregate pinned copy bandwidth with every GPU copying at once
mkdir -p gpurun_out
{ nvidia-smi topo -m; echo; lscpu | grep -i -E "model name|socket|numa|^cpu\(s\)|thread"; echo; numactl -H 2>/dev/null || cat /sys/devices/system/node/online; echo; free -g | head -2; } > gpurun_out/s19_topo.txt 2>&1
cat gpurun_out/s19_topo.txt | cut -c1-160
timeout 300 python tools/pcie_aggregate.py --gpus 8 --seconds 2 > gpurun_out/s19_pcie.json 2> gpurun_out/s19_pcie.err; cat gpurun_out/s19_pcie.json; tail -2 gpurun_out/s19_pcie.err
