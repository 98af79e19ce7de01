# coalesced host seam under concurrent callers: "threads pinned cpus [ENV=VAL ...]" per line (diagnosis; not a benchmark line)
export CSBWA_CO_TIMING=1
while read -r t p c extra; do
[ -z "$t" ] && continue
echo "== threads $t pinned $p cpus $c $extra"
env $extra timeout 300 python tools/e2e_probe.py --pairs 250000 --threads $t --pinned $p --cpus $c --repeat ${PROBE_REPEAT:-60} 2>&1 | grep -v Warning | tail -4
done <<CFG
${PROBE_CFGS:-64 1 0}
CFG
