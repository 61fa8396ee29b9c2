# speed of the precision classes on the batch-1 and batch-256 inference workloads
for P in fp16x3; do for WL in infer256_b1 infer256 infer1080_b1; do
python bench.py --workload $WL --precision $P --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bench_${WL}_$P.json 2> gpurun_out/err.txt || tail -3 gpurun_out/err.txt
python -c "
import json; d=json.load(open('gpurun_out/bench_${WL}_$P.json')); print('$P', '$WL', round(d['value'],1), d['unit'], round(d['ms_per_step'],3), 'ms/step; dominant-kernel', round(d['roofline']['kernel_ms']*1e3,1), 'us', 'algorithmic TFLOP/s', round(d['roofline']['achieved'],1))"
done; done
