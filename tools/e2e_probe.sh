# coalesced host seam under concurrent callers: "threads slots sync [ENV=VAL ...]" per line (diagnosis; not a benchmark line)
export CSBWA_CO_TIMING=1
while read -r t s m extra; do
[ -z "$t" ] && continue
echo "== threads $t slots $s sync $m $extra"
env CSBWA_CO_SYNC=$m CSBWA_CO_SLOTS=$s $extra timeout 200 python tools/e2e_probe.py --pairs 250000 --threads $t --repeat 40 2>&1 | grep -v Warning | tail -2
done <<CFG
${PROBE_CFGS:-64 16 sleep}
CFG
