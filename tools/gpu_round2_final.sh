#!/bin/bash
# Round 2 evidence bundle (one gpurun call, one GPU): default bench line, fp16x3 training line, reference arm, real timeline,
# kernel micro-benchmarks, ncu launch list of the training step and ncu --set full captures of the kernels BASELINE names
# (each ncu command runs to completion without ncu first; numbers printed under ncu are never bench values).
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "default rc=$?"
timeout 300 python bench.py --workload train --precision fp16x3 --steps 20 --warmup 5 > gpurun_out/r02_bench_train_fp16x3.json 2> gpurun_out/r02_bench_train_fp16x3.err; echo "x3 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_train.json 2> gpurun_out/r02_bench_reference_train.err; echo "reference rc=$?"
timeout 300 python bench.py --workload infer1080 --no-cpu-baseline --steps 8 > gpurun_out/r02_bench_infer1080.json 2> gpurun_out/r02_bench_infer1080.err; echo "infer1080 rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_default.json'))
print('train', round(d['ms_per_step'], 3), 'ms', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'roofline', round(d['roofline']['achieved']), round(d['roofline']['frac'], 3))
print('norm', {k: (round(v['achieved']), round(v['frac'], 3), round(v['us_per_launch'], 2)) for k, v in d['roofline_norm'].items() if k.endswith('_case')})
for k, v in d['inference'].items():
    print(k, round(v['value'], 1), 'e2e', round(v['e2e']['value'], 1), 'u8', round(v['e2e_uint8']['value'], 1), 'roof', round(v['roofline']['frac'], 3))
print('eager', json.dumps(d.get('gpu_eager_reference'))[:700])
x = json.load(open('gpurun_out/r02_bench_train_fp16x3.json')); print('train fp16x3', round(x['ms_per_step'], 3), 'ms', round(x['value'], 1))
PY
timeout 300 python tools/prof_train_timeline.py 4 > gpurun_out/r02_timeline.json 2> gpurun_out/r02_timeline.err; rm -f gpurun_out/train_trace.json; echo "timeline rc=$?"
timeout 300 python tools/bench_kernels.py --out gpurun_out/r02_bench_kernels.json > gpurun_out/r02_bench_kernels.log 2>&1; echo "kernels rc=$?"
timeout 120 python tools/bench_rowconv.py > gpurun_out/r02_bench_rowconv.json 2> gpurun_out/r02_bench_rowconv.err; echo "rowconv rc=$?"
timeout 120 python tools/exp_rowconv_timeline.py > gpurun_out/r02_rowconv_timeline.log 2>&1; echo "rowconv timeline rc=$?"
timeout 200 python tools/exp_conv_timeline.py > gpurun_out/r02_conv_timeline.log 2>&1; echo "conv timeline rc=$?"
export FNST_BENCH_NO_ROOFLINE=1
CMD="python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r02_plain_train.log 2>&1 && {
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 800 --csv --log-file gpurun_out/r02_launches_train.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1; echo "ncu list rc=$?"
for K in "^conv_tc_kernel" "rowconv_tc_kernel" "wgrad_tc_kernel" "inorm_apply_kernel|inorm_bwd_reduce_kernel|inorm_bwd_apply_kernel"; do
  N=$(echo $K | cut -d'|' -f1 | tr -d '^')
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 12 -c 6 -f -o gpurun_out/r02_prof_$N $CMD > gpurun_out/r02_ncu_full_$N.log 2>&1; echo "ncu full $N rc=$?"
done
}
python tools/summarize_launches.py gpurun_out/r02_launches_train.csv "Round 2: launches 1700..2500 of FNST_BENCH_NO_ROOFLINE=1 bench.py --workload train --steps 3 --warmup 3" > gpurun_out/r02_train_launches.md 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02_prof_*.ncu-rep > gpurun_out/r02_ncu_train_kernels.json 2> gpurun_out/r02_ncu_summary.err
find gpurun_out -name '*.ncu-rep' -size +12M -delete
du -sh gpurun_out | tail -1
