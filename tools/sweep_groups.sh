mkdir -p gpurun_out
for m in 0 1; do for g in 32 64 128 245; do for s in 2 4; do
timeout 300 python bench.py --ext-mode $m --group-calls $g --streams $s --steps 3 --no-cpu-baseline > gpurun_out/sw_${m}_${g}_${s}.json 2>/dev/null
python -c "
import json,sys
d=json.load(open('gpurun_out/sw_${m}_${g}_${s}.json')); print('mode',$m,'group',$g,'streams',$s,'value %.1f e2e %.1f frac %.3f'%(d['value'], d['e2e']['value'], d['roofline']['frac']), d['roofline']['phase_ms_sample'])
"
done; done; done
