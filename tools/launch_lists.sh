#!/bin/bash
# ncu launch lists of the bench commands (roofline leg skipped so that the list shows the step itself)
export FNST_BENCH_NO_ROOFLINE=1
for WL in train infer256; do
CMD="python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$WL.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_$WL.csv $CMD > gpurun_out/ncu_$WL.log 2>&1; echo "ncu $WL rc=$?"
done
