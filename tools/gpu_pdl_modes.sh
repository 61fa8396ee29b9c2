#!/bin/bash
# experiment: where to issue griddepcontrol.launch_dependents (libfnst_m2 = -DFNST_PDL_MODE=2 build: never explicitly)
mkdir -p gpurun_out
P=$PWD/fast_neural_style_transfer_b200
for V in off m3 m2; do
  case $V in off) L=$P/libfnst.so; PDL=0;; m3) L=$P/libfnst.so; PDL=1;; m2) L=$P/libfnst_m2.so; PDL=1;; esac
  for WL in train infer256_b1 infer256; do
    FNST_LIB=$L FNST_PDL=$PDL FNST_BENCH_NO_ROOFLINE=1 timeout 600 python bench.py --workload $WL --no-cpu-baseline --steps 30 > gpurun_out/mode_${WL}_$V.json 2> gpurun_out/mode_${WL}_$V.err
    python -c "
import json; d=json.load(open('gpurun_out/mode_${WL}_$V.json')); print('$V $WL', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value'],1))"
  done
done
