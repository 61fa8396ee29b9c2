#!/bin/bash
# usage: gpu_ab_env.sh VAR "v1 v2 ..." "workloads"  -- bench.py A/B over one environment knob
VAR=$1; VALS=$2; WLS=${3:-"train infer256_b1 infer256"}
mkdir -p gpurun_out
for R in 1 2; do
for V in $VALS; do
  for WL in $WLS; do
    env $VAR=$V FNST_BENCH_NO_ROOFLINE=1 timeout 600 python bench.py --workload $WL --no-cpu-baseline --steps 30 > gpurun_out/ab_${VAR}_${V}_$WL.json 2> gpurun_out/ab_${VAR}_${V}_$WL.err
    python -c "
import json; d=json.load(open('gpurun_out/ab_${VAR}_${V}_$WL.json')); print('round $R $VAR=$V $WL', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value'],1), d['clocks'])"
  done
done
done
