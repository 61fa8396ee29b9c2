#!/bin/bash
# Round 2 (gpurun --gpus 8): weak scaling of the training step at N = 1 and 8 in the same box, with the single post-backward
# all-reduce (default) and with the bucket-wise exchange under the staged backward (FNST_DP_OVERLAP=1); inference sharding.
mkdir -p gpurun_out
export FNST_BENCH_NO_ROOFLINE=1
run() {  # name, nproc, extra env...
  local name=$1 n=$2; shift 2
  if [ "$n" = 1 ]; then
    env "$@" timeout 300 python bench.py --workload train --no-cpu-baseline --steps 60 > gpurun_out/r02_8gpu_$name.json 2> gpurun_out/r02_8gpu_$name.err
  else
    env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 200)) bench.py --gpus $n --workload train --no-cpu-baseline --steps 60 > gpurun_out/r02_8gpu_$name.json 2> gpurun_out/r02_8gpu_$name.err
  fi
  echo "$name rc=$?"
}
run train_n1 1 A=1
run train_n8 8 A=1
run train_n8_overlap 8 FNST_DP_OVERLAP=1
run train_n4 4 A=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29777 bench.py --gpus 8 --workload infer256 --no-cpu-baseline --steps 8 > gpurun_out/r02_8gpu_infer256_n8.json 2> gpurun_out/r02_8gpu_infer256_n8.err; echo "infer256 n8 rc=$?"
python - <<'PY'
import json
for n in ("train_n1", "train_n4", "train_n8", "train_n8_overlap", "infer256_n8"):
    try:
        d = json.loads(open(f'gpurun_out/r02_8gpu_{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'], 4), 'ms value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1))
    except Exception as e:
        print(n, 'ERR', e); print(open(f'gpurun_out/r02_8gpu_{n}.err').read()[-600:])
PY
