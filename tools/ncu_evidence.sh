#!/bin/bash
# ncu --set full captures of the kernel classes named in BASELINE.json north_star (one gpurun call, ncu only).
# Each command first runs to completion without ncu.  Summaries (tools/ncu_summary.py) are produced on the box; reports larger
# than 15 MB are dropped so that gpurun_out/ stays below its 64 MiB limit.
mkdir -p gpurun_out
export FNST_BENCH_NO_ROOFLINE=1
CMD1="python bench.py --workload infer256 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD1 > gpurun_out/plain1.log 2>&1 && {
ncu --set full --clock-control none -k regex:"conv_tc_kernel|finalconv_tc_kernel|inorm_apply_kernel" -s 66 -c 4 -o gpurun_out/prof_infer256_trunk $CMD1 > gpurun_out/ncu1.log 2>&1; echo "ncu trunk rc=$?"
ncu --set full --clock-control none -k regex:"finalconv_tc_kernel" -s 2 -c 1 -o gpurun_out/prof_infer256_final $CMD1 > gpurun_out/ncu1b.log 2>&1; echo "ncu final rc=$?"
}
export FNST_CUDA_GRAPH=0
CMD2="python bench.py --workload train --steps 1 --warmup 3 --no-cpu-baseline"
$CMD2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none -k regex:"wgrad_tc_kernel|inorm_bwd" -s 60 -c 6 -o gpurun_out/prof_train_kernels $CMD2 > gpurun_out/ncu2.log 2>&1; echo "ncu train rc=$?"
for R in prof_infer256_trunk prof_infer256_final prof_train_kernels; do
  [ -f gpurun_out/$R.ncu-rep ] && python tools/ncu_summary.py gpurun_out/$R.ncu-rep > gpurun_out/$R.json 2> gpurun_out/$R.err
done
ls -la gpurun_out/*.ncu-rep
find gpurun_out -name '*.ncu-rep' -size +15M -delete
du -sh gpurun_out
