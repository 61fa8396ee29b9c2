python tools/bench_apply.py
timeout 900 python -m pytest -q --timeout 300 -p no:cacheprovider tests -m gpu 2>&1 | tail -3
python bench.py --workload infer256 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
