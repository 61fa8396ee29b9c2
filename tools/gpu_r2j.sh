#!/bin/bash
# Round 2, call J: the whole GPU suite (incl. reference scripts end to end on the drop-in).
mkdir -p gpurun_out
timeout 2400 python -m pytest -q --timeout 900 --timeout-method thread -p no:cacheprovider tests -m gpu -x > gpurun_out/r2j_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -25 gpurun_out/r2j_gpu_suite.log
timeout 600 python -m pytest -q -p no:cacheprovider tests/test_gpu_reference_scripts.py -m gpu -s 2>&1 | grep -E "rel_l2|diff|Iter \[|passed|failed" | tail -15
