python -m pytest tests -m gpu -x -q 2>&1 | tail -1
python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], {k: round(v['ms_per_step'],1) for k,v in d['kernels'].items()})"
python bench.py --records 300 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('R=300', d['ms_per_step'], {k: round(v['ms_per_step'],1) for k,v in d['kernels'].items()})"
