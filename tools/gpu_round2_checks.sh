#!/bin/bash
# Round 2 regression bundle for one gpurun call: the whole GPU suite, the default bench line, an A/B of one environment knob.
#   gpurun --timeout 3000 -- 'bash tools/gpu_round2_checks.sh [KNOB=VALUE ...]'
mkdir -p gpurun_out
timeout 2400 python -m pytest -q --timeout 900 --timeout-method thread -p no:cacheprovider tests -m gpu > gpurun_out/r02_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/r02_gpu_suite.log
run() { local name=$1; shift; env FNST_BENCH_NO_ROOFLINE=1 "$@" timeout 300 python bench.py --workload train --no-cpu-baseline --steps 40 > gpurun_out/r02_ab_$name.json 2> gpurun_out/r02_ab_$name.err; python -c "
import json; d=json.load(open('gpurun_out/r02_ab_$name.json')); print('$name', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value'],1), d['gpu_launches'])" || tail -3 gpurun_out/r02_ab_$name.err; }
run default
for kv in "$@"; do run "${kv//=/_}" "$kv"; done
