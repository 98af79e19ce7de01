# resident leg of bench.py against streams / calls per launch sequence (tuning aid; not a benchmark line)
mkdir -p gpurun_out
for cfg in ${SWEEP_CFGS:-"32:4"}; do
g=${cfg%%:*}; s=${cfg##*:}
timeout 300 python bench.py --group-calls $g --streams $s --steps 5 --no-cpu-baseline --no-matesw --no-e2e > gpurun_out/sw_${g}_${s}.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/sw_${g}_${s}.json')); print('group',$g,'streams',$s,'value %.1f frac %.3f'%(d['value'], d['roofline']['frac']))
"
done
