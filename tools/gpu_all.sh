#!/bin/bash
# full GPU test-suite + default bench (+ optional extra workloads as arguments)
mkdir -p gpurun_out
timeout 1500 python -m pytest -q --timeout 300 --timeout-method thread -p no:cacheprovider tests -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for WL in train "$@"; do
python bench.py --workload $WL --no-cpu-baseline > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err; echo "$WL rc=$?"; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$WL.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])"; tail -2 gpurun_out/bench_$WL.err
done
