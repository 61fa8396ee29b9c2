mkdir -p gpurun_out
for WL in train infer256; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $WL --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_$WL.json 2> gpurun_out/bench2_$WL.err; echo "2gpu $WL rc=$?"; tail -1 gpurun_out/bench2_$WL.json; tail -3 gpurun_out/bench2_$WL.err
done
python bench.py --gpus 1 --workload train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench1_train.json 2>/dev/null; tail -1 gpurun_out/bench1_train.json
python bench.py --impl reference --workload train --steps 2 --warmup 1 | tail -1
