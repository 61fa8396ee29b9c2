#!/bin/bash
# Round 2, call B: new operators (one-pass InstanceNorm backward, bf16 twins, loss scalars), training parity, step timing.
mkdir -p gpurun_out
timeout 900 python -m pytest -q -x --timeout 300 --timeout-method thread -p no:cacheprovider tests/test_gpu_bwd_ops.py tests/test_gpu_ops.py -m gpu > gpurun_out/r2b_ops.log 2>&1; echo "ops rc=$?"; tail -15 gpurun_out/r2b_ops.log
timeout 900 python -m pytest -q --timeout 600 --timeout-method thread -p no:cacheprovider tests/test_gpu_train.py tests/test_gpu_dropin.py tests/test_gpu_baseline_configs.py -m gpu > gpurun_out/r2b_train.log 2>&1; echo "train rc=$?"; tail -15 gpurun_out/r2b_train.log
timeout 300 python tools/prof_train_parts.py > gpurun_out/r2b_parts.log 2>&1; echo "parts rc=$?"; tail -16 gpurun_out/r2b_parts.log
FNST_BENCH_NO_ROOFLINE=1 timeout 300 python bench.py --workload train --no-cpu-baseline --steps 40 > gpurun_out/r2b_bench_train.json 2> gpurun_out/r2b_bench_train.err; echo "bench rc=$?"; cat gpurun_out/r2b_bench_train.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'])"; tail -3 gpurun_out/r2b_bench_train.err
