#!/bin/bash
# One gpurun call: operator + network parity on the B200, logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
T="timeout 600 python -m pytest -q --timeout 180 --timeout-method thread -p no:cacheprovider"
$T tests/test_gpu_ops.py -m gpu -k "not conv_tc" > gpurun_out/ops_simt.log 2>&1; echo "ops_simt rc=$?"
$T tests/test_gpu_ops.py -m gpu -k "conv_tc" > gpurun_out/ops_tc.log 2>&1; echo "ops_tc rc=$?"
$T tests/test_gpu_net.py -m gpu -s > gpurun_out/net.log 2>&1; echo "net rc=$?"
tail -n 25 gpurun_out/ops_simt.log gpurun_out/ops_tc.log gpurun_out/net.log
