timeout 1200 python -m pytest -q --timeout 300 --timeout-method thread -p no:cacheprovider tests -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
for WL in train; do
python bench.py --workload $WL --no-cpu-baseline > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err; echo "$WL rc=$?"; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$WL.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['last_losses'])"; tail -2 gpurun_out/bench_$WL.err
done
