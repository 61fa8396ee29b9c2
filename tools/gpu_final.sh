#!/bin/bash
# Final round-1 measurement pass: GPU tests, every bench workload (default = train with cpu_baseline and roofline legs),
# the reference arm, one `ncu --set full` capture of the optimizer-tail / InstanceNorm-backward kernels, and a launch list
# of two training steps.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
timeout 300 python -m pytest -q --timeout 300 --timeout-method thread -p no:cacheprovider tests -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default rc=$?"
for WL in infer256 infer1080_b1 infer1080 infer256_b1; do
  timeout 150 python bench.py --workload $WL > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err; echo "$WL rc=$?"
done
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_train.json 2> gpurun_out/bench_reference_train.err; echo "reference rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], round(d['value'], 1), d['unit'], 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1),
              'frac', d.get('roofline', {}).get('frac'), 'tail', d.get('roofline_optimizer_tail', {}).get('frac'), 'cpu', d.get('cpu_baseline', {}).get('value'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
FNST_BENCH_NO_ROOFLINE=1 timeout 150 ncu --set full --clock-control none --import-source on -k regex:'mt_|inorm_bwd_reduce' --launch-skip 30 -c 8 \
  -o gpurun_out/prof_optim_tail -f python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_optim.log 2>&1; echo "ncu full rc=$?"
FNST_BENCH_NO_ROOFLINE=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2600 -c 700 --csv \
  --log-file gpurun_out/launches_train_r1c.csv python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train_r1c.log 2>&1; echo "ncu list rc=$?"
