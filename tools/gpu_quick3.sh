#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -q --timeout 300 --timeout-method thread -p no:cacheprovider tests -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for WL in train infer256_b1 infer256 infer1080 infer1080_b1; do
  FNST_BENCH_NO_ROOFLINE=1 timeout 600 python bench.py --workload $WL --no-cpu-baseline --steps 30 > gpurun_out/q_$WL.json 2> gpurun_out/q_$WL.err
  python -c "
import json; d=json.load(open('gpurun_out/q_$WL.json')); print('$WL', round(d['ms_per_step'],4), 'ms  value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'])"
done
