mkdir -p gpurun_out
CMD="python bench.py --workload infer256 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_full.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 14 -c 3 -o gpurun_out/prof_conv_tc $CMD > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
for WL in train infer256; do
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$WL.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$WL.csv $CMD > gpurun_out/ncu_$WL.log 2>&1; echo "ncu $WL rc=$?"
done
