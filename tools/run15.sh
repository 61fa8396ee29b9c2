CMD="python bench.py --workload train --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_train.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 3500 -c 2000 --csv --log-file gpurun_out/launches_train.csv $CMD > gpurun_out/ncu_train.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_train.log | cut -c1-300
