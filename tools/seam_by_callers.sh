#!/bin/bash
# End-to-end throughput and latency of the extension seam by number of blocking callers (4096-read calls of C2, pinned
# caller buffers), with the small-group lane-group kernel on (default) and off.  One JSON line per run:
#   bash tools/seam_by_callers.sh > profiles/r2_seam_by_callers.jsonl
cd "$(dirname "$0")/.."
for t in 1 2 3 4 8 16 32 64; do
  for coop in 8192 0; do
    CSBWA_EXT_COOP_MAX=$coop timeout 300 python tools/e2e_probe.py --pairs 250000 --threads $t --pinned 1 --repeat ${PROBE_REPEAT:-20} 2>&1 | grep '^{' | tail -1
  done
done
