#!/bin/bash
# session 16: lane-group kernel as a THROUGHPUT path (measured side by side with the class kernels),
# one ncu capture of it, and the seam's busy policy
cd /root/repo
mkdir -p gpurun_out
B="python bench.py --pairs 250000 --steps 3 --warmup 3 --no-e2e --no-matesw --no-cpu-baseline"
: > gpurun_out/s16_lane_group.jsonl
for wl in C5 C2; do
  echo "{\"run\": \"$wl class kernels (thread per side)\"}" >> gpurun_out/s16_lane_group.jsonl
  CSBWA_EXT_COOP_MAX=0 $B --workload $wl 2>/dev/null | tail -1 >> gpurun_out/s16_lane_group.jsonl
  for g in 8 16 32; do
    echo "{\"run\": \"$wl lane-group kernel, $g lanes per task\"}" >> gpurun_out/s16_lane_group.jsonl
    CSBWA_EXT_COOP_MAX=1000000000 CSBWA_EXT_COOP_G=$g $B --workload $wl 2>/dev/null | tail -1 >> gpurun_out/s16_lane_group.jsonl
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/s16_lane_group.jsonl"):
    d = json.loads(l)
    if "run" in d: print(d["run"], end=": ")
    else: print(d.get("value"), d.get("unit"))
PY
# ncu: the lane-group kernel on C5 (8 lanes), plain launches
CSBWA_EXT_COOP_MAX=1000000000 CSBWA_EXT_COOP_G=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ext_small -c 2 \
  -o gpurun_out/s16_ext_small_g8 $B --workload C5 --pairs 60000 --no-graph --steps 1 --warmup 1 > gpurun_out/s16_ncu.log 2>&1
CSBWA_EXT_COOP_MAX=1000000000 CSBWA_EXT_COOP_G=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ext_small -c 2 \
  -o gpurun_out/s16_ext_small_g32 $B --workload C2 --pairs 60000 --no-graph --steps 1 --warmup 1 >> gpurun_out/s16_ncu.log 2>&1
tail -3 gpurun_out/s16_ncu.log
export PROBE_REPEAT=20
PROBE_CFGS="2 1 0 CSBWA_EXT_COOP_BUSY=2
3 1 0 CSBWA_EXT_COOP_BUSY=2
4 1 0 CSBWA_EXT_COOP_BUSY=2
4 1 0 CSBWA_EXT_COOP_BUSY=3
8 1 0 CSBWA_EXT_COOP_BUSY=2
16 1 0 CSBWA_EXT_COOP_BUSY=2
16 1 0 CSBWA_EXT_COOP_BUSY=1" bash tools/e2e_probe.sh > gpurun_out/s16_probe.log 2>&1
grep -E "^==|gcups" gpurun_out/s16_probe.log | sed -E 's/.*"gcups": ([0-9.]+), "calls_per_group": ([0-9.]+), "ms_per_group": \{"host_ms": ([0-9.]+).*/  gcups \1 calls\/group \2 ms\/group \3/'
