#!/bin/bash
# Round-1 late check: GPU tests, optimizer-tail A/B (libfnst multi-tensor kernels vs torch foreach), InstanceNorm-backward
# occupancy A/B, kernel micro-benchmarks of both.
mkdir -p gpurun_out
timeout 600 python -m pytest -q --timeout 300 --timeout-method thread -p no:cacheprovider tests -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
run() {  # name, extra env / args
  local name=$1; shift
  env FNST_BENCH_NO_ROOFLINE=1 "$@" > gpurun_out/q_$name.json 2> gpurun_out/q_$name.err; echo "$name rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/q_$name.json')); print('$name', round(d['ms_per_step'],4), 'ms  value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'])" || tail -3 gpurun_out/q_$name.err
}
run train_fnst timeout 300 python bench.py --workload train --no-cpu-baseline --steps 40 --optimizer fnst
run train_torch timeout 300 python bench.py --workload train --no-cpu-baseline --steps 40 --optimizer torch
run train_fnst_b1 FNST_INORM_BWD_BLOCKS=1 timeout 300 python bench.py --workload train --no-cpu-baseline --steps 40 --optimizer fnst
timeout 200 python tools/bench_kernels.py --only inorm_bwd --out gpurun_out/bench_kernels_inorm_bwd.json 2>&1 | tail -8
timeout 200 python tools/bench_kernels.py --only optimizer_tail --out gpurun_out/bench_kernels_optim.json 2>&1 | tail -6
