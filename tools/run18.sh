python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default rc=$?"; cat gpurun_out/bench_default.json | cut -c1-3000
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-600
for WL in infer256 infer1080 infer1080_b1 infer256_b1; do python bench.py --workload $WL > gpurun_out/bench_$WL.json 2>gpurun_out/bench_$WL.err; echo "$WL rc=$?"; done
