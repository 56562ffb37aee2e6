timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest.log 2>&1; tail -2 gpurun_out/r02_final_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline --steps 2 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
