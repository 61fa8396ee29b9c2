#!/bin/bash
# Round 2 (gpurun --gpus 2): NCCL 2-rank parity test; data-parallel bench at N=2 (bucket-wise exchange under the backward, and
# the single post-backward all-reduce for comparison) next to N=1 in the same box.
mkdir -p gpurun_out
timeout 900 python -m pytest -q --timeout 600 --timeout-method thread -p no:cacheprovider tests/test_gpu_parallel_nccl.py -m gpu -s > gpurun_out/r02_2gpu_nccl.log 2>&1; echo "nccl test rc=$?"; grep -E "NCCL gradient|passed|failed|Error" gpurun_out/r02_2gpu_nccl.log | tail
export FNST_BENCH_NO_ROOFLINE=1
timeout 300 python bench.py --workload train --no-cpu-baseline --steps 60 > gpurun_out/r02_2gpu_train_n1.json 2> gpurun_out/r02_2gpu_train_n1.err; echo "n1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload train --no-cpu-baseline --steps 60 > gpurun_out/r02_2gpu_train_n2.json 2> gpurun_out/r02_2gpu_train_n2.err; echo "n2 rc=$?"
FNST_DP_OVERLAP=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload train --no-cpu-baseline --steps 60 > gpurun_out/r02_2gpu_train_n2_overlap.json 2> gpurun_out/r02_2gpu_train_n2_overlap.err; echo "n2 (bucketed overlap) rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload infer256 --no-cpu-baseline --steps 8 > gpurun_out/r02_2gpu_infer256_n2.json 2> gpurun_out/r02_2gpu_infer256_n2.err; echo "infer n2 rc=$?"
python - <<'PY'
import json
for n in ("train_n1","train_n2","train_n2_overlap","infer256_n2"):
    try:
        d=json.loads(open(f'gpurun_out/r02_2gpu_{n}.json').read().strip().splitlines()[-1]); print(n, round(d['ms_per_step'],4), 'ms value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))
    except Exception as e: print(n, 'ERR', e); print(open(f'gpurun_out/r02_2gpu_{n}.err').read()[-600:])
PY
