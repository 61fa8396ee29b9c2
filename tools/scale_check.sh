#!/bin/bash
# multi-GPU check: bash tools/scale_check.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
for WL in train infer256; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $WL --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_${WL}_$N.json 2> gpurun_out/scale_${WL}_$N.err; echo "N=$N $WL rc=$?"; python -c "
import json,sys; d=json.loads(open('gpurun_out/scale_${WL}_$N.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'])"; grep -iE "error|Traceback" gpurun_out/scale_${WL}_$N.err | head -3
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --impl reference --workload train --steps 1 --warmup 1 | cut -c1-200
