#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest -q --timeout 900 --timeout-method thread -p no:cacheprovider tests/test_gpu_train.py tests/test_gpu_baseline_configs.py tests/test_gpu_dropin.py -m gpu -s > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"; grep -E "worst|global|passed|failed|FAILED|Error|rel_l2" gpurun_out/r2h_tests.log | tail -40
for prec in fp16 fp16x3; do FNST_BENCH_NO_ROOFLINE=1 timeout 300 python bench.py --workload train --precision $prec --no-cpu-baseline --steps 40 > gpurun_out/r2h_train_$prec.json 2> gpurun_out/r2h_train_$prec.err; python -c "
import json; d=json.load(open('gpurun_out/r2h_train_$prec.json')); print('$prec', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value'],1), d['gpu_launches'])" || tail -3 gpurun_out/r2h_train_$prec.err; done
