#!/bin/bash
# ncu --set full captures of the kernel classes named in BASELINE.json north_star (one gpurun call, ncu only)
mkdir -p gpurun_out
CMD1="python bench.py --workload infer256 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD1 > gpurun_out/plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:inorm_apply_kernel -s 20 -c 4 -o gpurun_out/prof_inorm_apply $CMD1 > gpurun_out/ncu1.log 2>&1; echo "ncu apply rc=$?"
export FNST_CUDA_GRAPH=0
CMD2="python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline"
$CMD2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel|inorm_bwd" -s 60 -c 12 -o gpurun_out/prof_train_kernels $CMD2 > gpurun_out/ncu2.log 2>&1; echo "ncu train rc=$?"
tail -2 gpurun_out/ncu1.log gpurun_out/ncu2.log | cut -c1-200
