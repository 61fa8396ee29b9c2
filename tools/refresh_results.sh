#!/bin/bash
# refresh profiles/r01_bench_*.json with the current build (1 GPU)
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default rc=$?"
for WL in infer256 infer1080 infer1080_b1 infer256_b1; do python bench.py --workload $WL --steps 20 > gpurun_out/bench_$WL.json 2>gpurun_out/bench_$WL.err; echo "$WL rc=$?"; done
python bench.py --workload infer256_b1 --precision fp16x3 --steps 20 --no-cpu-baseline > gpurun_out/bench_infer256_b1_fp16x3.json 2>/dev/null; echo "x3 rc=$?"
python bench.py --workload infer256_b1 --precision fp32 --steps 20 --no-cpu-baseline > gpurun_out/bench_infer256_b1_fp32.json 2>/dev/null; echo "fp32 rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_train.json; echo "ref rc=$?"
bash tools/launch_lists.sh
