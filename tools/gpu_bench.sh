#!/bin/bash
# bench + ncu launch list of the same command (B200, 1 GPU)
mkdir -p gpurun_out
WL=${1:-infer256}
python bench.py --workload $WL > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err; echo "bench rc=$?"; cat gpurun_out/bench_$WL.json; tail -5 gpurun_out/bench_$WL.err
python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$WL.csv \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu.log
