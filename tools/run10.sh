for WL in infer256_b1 infer1080_b1 infer1080; do
python bench.py --workload $WL --steps 20 --warmup 5 > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err; echo "$WL rc=$?"; cut -c1-1400 gpurun_out/bench_$WL.json; tail -2 gpurun_out/bench_$WL.err
done
