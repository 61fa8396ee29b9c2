timeout 1200 python -m pytest -q --timeout 300 --timeout-method thread -p no:cacheprovider tests -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
bash tools/gpu_bench.sh train
