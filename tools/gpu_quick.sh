#!/bin/bash
# GPU regression + quick numbers: full -m gpu suite, CTA-pair check, conv fixed-cost experiment, 3 benches
mkdir -p gpurun_out
timeout 900 python -m pytest -q --timeout 300 --timeout-method thread -p no:cacheprovider tests -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 100 python tools/check_pair.py > gpurun_out/check_pair.log 2>&1; echo "check_pair rc=$?"; head -8 gpurun_out/check_pair.log
timeout 100 python tools/exp_conv_fixed_cost.py > gpurun_out/exp_conv_fixed_cost.log 2>&1; head -12 gpurun_out/exp_conv_fixed_cost.log
for WL in train infer256_b1 infer256; do
  FNST_BENCH_NO_ROOFLINE=1 timeout 600 python bench.py --workload $WL --no-cpu-baseline --steps 30 > gpurun_out/q_$WL.json 2> gpurun_out/q_$WL.err
  python -c "
import json; d=json.load(open('gpurun_out/q_$WL.json')); print('$WL', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value'],1), d['clocks'])"
done
