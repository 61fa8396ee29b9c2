#!/bin/bash
# Round 2, call C: why is the step slower?  A/B of the one-pass InstanceNorm backward, kernel micro-benchmarks, launch list.
mkdir -p gpurun_out
timeout 600 python -m pytest -q -x --timeout 300 --timeout-method thread -p no:cacheprovider tests/test_gpu_bwd_ops.py tests/test_gpu_ops.py -m gpu > gpurun_out/r2c_ops.log 2>&1; echo "ops rc=$?"; tail -5 gpurun_out/r2c_ops.log
for only in inorm_bwd grad_assemble; do timeout 300 python tools/bench_kernels.py --only $only --out gpurun_out/r2c_bk_$only.json 2>&1 | tail -12; done
run() { local name=$1; shift; env FNST_BENCH_NO_ROOFLINE=1 "$@" timeout 300 python bench.py --workload train --no-cpu-baseline --steps 40 > gpurun_out/r2c_$name.json 2> gpurun_out/r2c_$name.err; echo "$name rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2c_$name.json')); print('$name', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value'],1), d['gpu_launches'])" || tail -3 gpurun_out/r2c_$name.err; }
run fused1 FNST_INORM_BWD_FUSED=1
run fused0 FNST_INORM_BWD_FUSED=0
run nograph FNST_CUDA_GRAPH=0
FNST_BENCH_NO_ROOFLINE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 800 --csv --log-file gpurun_out/r2c_launches_train.csv python bench.py --workload train --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2c_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r2c_launches_train.csv 2>/dev/null | head -60
