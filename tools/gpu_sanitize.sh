#!/bin/bash
# compute-sanitizer pass over small shapes: memcheck on the native resize check and on a slice of the
# operator tests, racecheck on the shared-memory-heavy kernels.  One gpurun call; slow (each tool serialises kernels).
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 300 $CS --tool memcheck --error-exitcode 9 tests/native/bin/resize_selftest > gpurun_out/sanitize_resize_memcheck.log 2>&1; echo "resize memcheck rc=$?"
timeout 300 $CS --tool racecheck --error-exitcode 9 tests/native/bin/resize_selftest > gpurun_out/sanitize_resize_racecheck.log 2>&1; echo "resize racecheck rc=$?"
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest -q -x -p no:cacheprovider tests/test_gpu_optim.py tests/test_gpu_preprocess.py \
  tests/test_gpu_bwd_ops.py -k "not native" > gpurun_out/sanitize_pytest_memcheck.log 2>&1; echo "pytest memcheck rc=$?"
tail -3 gpurun_out/sanitize_*.log
