#!/bin/bash
# Round 2, call K: staged (smem + TMA store) conv epilogue: parity, timeline, fixed cost, step A/B; pinned-host pipeline.
mkdir -p gpurun_out
timeout 900 python -m pytest -q -x --timeout 600 --timeout-method thread -p no:cacheprovider tests/test_gpu_ops.py tests/test_gpu_b200_features.py tests/test_gpu_net.py tests/test_gpu_dropin.py -m gpu > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2k_tests.log
for so in 1 0; do echo "stage_out=$so"; FNST_CONV_STAGE_OUT=$so timeout 200 python tools/exp_conv_timeline.py 2>&1 | tail -5; done
run() { local name=$1; shift; env FNST_BENCH_NO_ROOFLINE=1 "$@" timeout 300 python bench.py --workload train --no-cpu-baseline --steps 40 > gpurun_out/r2k_$name.json 2> gpurun_out/r2k_$name.err; echo "$name rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2k_$name.json')); print('$name', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value'],1), d['gpu_launches'])" || tail -3 gpurun_out/r2k_$name.err; }
run stage1 FNST_CONV_STAGE_OUT=1
run stage0 FNST_CONV_STAGE_OUT=0
timeout 300 python bench.py --workload infer256 --no-cpu-baseline --steps 8 > gpurun_out/r2k_infer256.json 2> gpurun_out/r2k_infer256.err; python -c "
import json; d=json.load(open('gpurun_out/r2k_infer256.json')); print('infer256', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'single-stream', round(d['e2e']['single_stream_value'],1), 'roof', round(d['roofline']['frac'],3))" || tail -3 gpurun_out/r2k_infer256.err
timeout 300 python bench.py --workload infer256_b1 --no-cpu-baseline --steps 50 > gpurun_out/r2k_infer256_b1.json 2> gpurun_out/r2k_infer256_b1.err; python -c "
import json; d=json.load(open('gpurun_out/r2k_infer256_b1.json')); print('infer256_b1 fp16', round(d['value'],1), round(d['ms_per_step'],4), 'roof', round(d['roofline']['achieved'],1))" || tail -3 gpurun_out/r2k_infer256_b1.err
