python tools/bench_apply.py
timeout 600 python -m pytest -q --timeout 300 -p no:cacheprovider tests/test_gpu_ops.py tests/test_gpu_net.py -m gpu 2>&1 | tail -3
