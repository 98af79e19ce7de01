# coalesced host seam under concurrent callers: sync mode x slots x threads (diagnosis; not a benchmark line)
export CSBWA_CO_TIMING=1
for cfg in ${PROBE_CFGS:-"64 16 sleep" "64 16 sleep2" "64 20 sleep2" "64 12 sleep2" "128 16 sleep2" "16 16 sleep2" "16 16 spin"}; do
set -- $cfg
CSBWA_CO_SYNC=$3 CSBWA_CO_SLOTS=$2 timeout 200 python tools/e2e_probe.py --pairs 250000 --threads $1 --repeat 8 2>&1 | grep -v Warning | tail -2
done
