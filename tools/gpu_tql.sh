echo "== full"; timeout 300 python tools/time_solver.py 2>&1 | tail -1
echo "== small 2368"; timeout 300 python tools/time_solver.py 2368 2>&1 | tail -1
echo "== tiny 296"; timeout 300 python tools/time_solver.py 296 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tql_pytest.log 2>&1; tail -3 gpurun_out/r02_tql_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_tql_bench.log 2> gpurun_out/r02_tql_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02_tql_bench.log").read().strip().splitlines()[-1])
    print(d["metric"], d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print({k:(round(v["ms_per_step"],1)) for k,v in d["kernels"].items()})
except Exception as e: print("no line", e)
PY
