#!/bin/bash
# Round 2, call G: full default bench line (train + inference sub-objects + eager yardstick), dropin tests.
mkdir -p gpurun_out
timeout 900 python -m pytest -q --timeout 600 --timeout-method thread -p no:cacheprovider tests/test_gpu_dropin.py tests/test_gpu_train.py tests/test_gpu_net.py -m gpu > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2g_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/r2g_bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench_default.json'))
print('train', round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'roofline', round(d['roofline']['achieved']), round(d['roofline']['frac'],3), 'cpu', d.get('cpu_baseline',{}).get('value'))
print('norm', {k:(round(v['achieved']),round(v['frac'],3),round(v['us_per_launch'],2)) for k,v in d['roofline_norm'].items() if k.endswith('_case')})
for k,v in d['inference'].items(): print(k, round(v['value'],1), v['unit'], 'e2e', round(v['e2e']['value'],1), 'u8', round(v['e2e_uint8']['value'],1), 'roof', round(v['roofline']['frac'],3))
print('eager', json.dumps(d.get('gpu_eager_reference'))[:900])
PY
