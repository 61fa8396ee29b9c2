#!/bin/bash
# A/B of programmatic dependent launch + per-kernel timings; everything into gpurun_out/
mkdir -p gpurun_out
timeout 1200 python -m pytest -q --timeout 300 --timeout-method thread -p no:cacheprovider tests -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for PDL in 0 1; do
for WL in train infer256_b1; do
FNST_PDL=$PDL timeout 600 python bench.py --workload $WL --no-cpu-baseline > gpurun_out/ab_${WL}_pdl$PDL.json 2> gpurun_out/ab_${WL}_pdl$PDL.err; echo "$WL pdl=$PDL rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/ab_${WL}_pdl$PDL.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'kernel_ms', d['roofline']['kernel_ms'])"
done; done
timeout 600 python tools/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1; echo "bench_kernels rc=$?"; cat gpurun_out/bench_kernels.log | tail -60
