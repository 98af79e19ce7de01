#!/bin/bash
# session 29: seam throughput / latency by caller count on the final build; fresh ncu capture of the fused side kernel
cd /root/repo; mkdir -p gpurun_out
bash tools/seam_by_callers.sh > gpurun_out/r2_seam_by_callers.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/r2_seam_by_callers.jsonl"):
    d=json.loads(l); print(d["threads"], "callers, small-group bound", d["small_group_max_tasks"], ": gcups", round(d["gcups"],1), "ms/call", round(d["ms_per_call"],3), "calls/group", round(d["calls_per_group"],2))
PY
export CSBWA_PROFILE_STEP=1
CSBWA_EXT_FUSED_MAX=1000000000 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_ext_side -s 6 -c 3 -o gpurun_out/s29_ext_both_full python bench.py --pairs 131072 --steps 1 --no-graph --streams 1 --no-cpu-baseline --no-e2e --no-matesw > /dev/null 2> gpurun_out/s29_ncu.err; echo "ncu rc=$?"
