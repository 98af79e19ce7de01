# e2e leg of bench.py against caller threads / submission slots (diagnosis; not a benchmark line)
mkdir -p gpurun_out
for cfg in "32 8" "64 8" "32 16" "64 16" "128 16" "16 8"; do set -- $cfg
CSBWA_CO_SLOTS=$2 timeout 300 python bench.py --steps 3 --no-cpu-baseline --threads $1 > gpurun_out/e2e_$1_$2.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/e2e_$1_$2.json')); print('threads',$1,'slots',$2,'value %.0f e2e %.0f'%(d['value'], d['e2e']['value']), 'calls/sub %.2f ms/sub %.2f'%(d['e2e_calls_per_device_submission'], d['e2e_ms_per_device_submission']), d['e2e_device_ms_per_submission'], 'cpus', d['host_cpus'])
"
done
