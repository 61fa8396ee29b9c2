#!/bin/bash
# Round 2, call D: TMA-staged one-pass InstanceNorm backward: parity, micro-benchmark, step A/B.
mkdir -p gpurun_out
timeout 600 python -m pytest -q -x --timeout 300 --timeout-method thread -p no:cacheprovider tests/test_gpu_bwd_ops.py -m gpu > gpurun_out/r2d_ops.log 2>&1; echo "ops rc=$?"; tail -5 gpurun_out/r2d_ops.log
timeout 300 python tools/bench_kernels.py --only inorm_bwd --out gpurun_out/r2d_bk_inorm_bwd.json 2>&1 | tail -12
run() { local name=$1; shift; env FNST_BENCH_NO_ROOFLINE=1 "$@" timeout 300 python bench.py --workload train --no-cpu-baseline --steps 40 > gpurun_out/r2d_$name.json 2> gpurun_out/r2d_$name.err; echo "$name rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2d_$name.json')); print('$name', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value'],1), d['gpu_launches'])" || tail -3 gpurun_out/r2d_$name.err; }
run fused1 FNST_INORM_BWD_FUSED=1
run fused0 FNST_INORM_BWD_FUSED=0
timeout 600 python -m pytest -q --timeout 600 --timeout-method thread -p no:cacheprovider tests/test_gpu_train.py -m gpu > gpurun_out/r2d_train.log 2>&1; echo "train rc=$?"; tail -4 gpurun_out/r2d_train.log
