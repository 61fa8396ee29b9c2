#!/bin/bash
# compute-sanitizer pass over small shapes: memcheck on the native resize check and on the operator / feature
# tests that exercise every kernel family (gather-GEMM incl. CTA pairs, pixel-stream form and staged epilogue, row-streaming
# kernels, weight gradients, InstanceNorm forward / backward incl. the TMA-staged and cluster forms, optimizer tail, resize).
# One gpurun call; slow (the tool serialises kernels).  Large-shape cases are deselected.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 300 $CS --tool memcheck --error-exitcode 9 tests/native/bin/resize_selftest > gpurun_out/sanitize_resize_memcheck.log 2>&1; echo "resize memcheck rc=$?"
timeout 1500 $CS --tool memcheck --error-exitcode 9 python -m pytest -q -x -p no:cacheprovider tests/test_gpu_optim.py tests/test_gpu_preprocess.py \
  tests/test_gpu_bwd_ops.py tests/test_gpu_ops.py tests/test_gpu_b200_features.py \
  -k "not native and not 2000 and not 1080 and not shape3 and not 160 and not 256-256" > gpurun_out/sanitize_pytest_memcheck.log 2>&1; echo "pytest memcheck rc=$?"
tail -4 gpurun_out/sanitize_*.log
grep -c "ERROR SUMMARY: 0 errors" gpurun_out/sanitize_*.log
