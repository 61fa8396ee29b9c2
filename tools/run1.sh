timeout 900 python -m pytest -q --timeout 180 --timeout-method thread -p no:cacheprovider tests -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
bash tools/gpu_bench.sh infer256
