#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -q --timeout 600 --timeout-method thread -p no:cacheprovider tests/test_gpu_b200_features.py tests/test_gpu_dropin.py -m gpu > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2l_tests.log
for wl in infer256 infer1080; do timeout 300 python bench.py --workload $wl --no-cpu-baseline --steps 8 > gpurun_out/r2l_$wl.json 2> gpurun_out/r2l_$wl.err; python -c "
import json; d=json.load(open('gpurun_out/r2l_$wl.json')); print('$wl', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'single-stream', round(d['e2e']['single_stream_value'],1), 'u8', round(d['e2e_uint8']['value'],1), 'roof', round(d['roofline']['frac'],3))" || tail -3 gpurun_out/r2l_$wl.err; done
timeout 300 python bench.py --workload train --no-cpu-baseline --steps 40 > gpurun_out/r2l_train.json 2> gpurun_out/r2l_train.err; python -c "
import json; d=json.load(open('gpurun_out/r2l_train.json')); print('train', round(d['ms_per_step'],4), 'roof', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],3), 'kernel_ms', d['roofline']['kernel_ms'])" || tail -3 gpurun_out/r2l_train.err
