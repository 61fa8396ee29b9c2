for WL in train infer256 infer256_b1; do
python bench.py --workload $WL --no-cpu-baseline > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err; echo "$WL rc=$?"; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$WL.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline'])"; tail -2 gpurun_out/bench_$WL.err
done
