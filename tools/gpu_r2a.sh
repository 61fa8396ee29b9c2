#!/bin/bash
# Round 2, call A: parity at the BASELINE configurations, per-CTA timeline of the gather-GEMM, phase timing of the step.
mkdir -p gpurun_out
timeout 900 python -m pytest -q --timeout 600 --timeout-method thread -p no:cacheprovider tests/test_gpu_baseline_configs.py -m gpu -s > gpurun_out/r2a_parity.log 2>&1; echo "parity rc=$?"; grep -E "rel|loss|grad|passed|failed|Error|error" gpurun_out/r2a_parity.log | tail -60
timeout 200 python tools/exp_conv_timeline.py > gpurun_out/r2a_timeline.log 2>&1; echo "timeline rc=$?"; cat gpurun_out/r2a_timeline.log | tail -8
timeout 300 python tools/prof_train_parts.py > gpurun_out/r2a_parts.log 2>&1; echo "parts rc=$?"; tail -16 gpurun_out/r2a_parts.log
