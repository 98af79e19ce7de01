# e2e leg of bench.py against caller threads / submission slots (diagnosis; not a benchmark line)
mkdir -p gpurun_out
for cfg in ${E2E_CFGS:-"64 8"}; do
t=${cfg%%:*}; s=${cfg##*:}
CSBWA_CO_SLOTS=$s timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-matesw --threads $t > gpurun_out/e2e_${t}_$s.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/e2e_${t}_$s.json')); print('threads',$t,'slots',$s,'value %.0f e2e %.0f'%(d['value'], d['e2e']['value']), 'calls/sub %.2f ms/sub %.2f'%(d['e2e_calls_per_device_submission'], d['e2e_ms_per_device_submission']), d['e2e_device_ms_per_submission'])
"
done
