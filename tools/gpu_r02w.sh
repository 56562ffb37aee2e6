echo "== wave4"; timeout 300 python tools/time_solver.py 2>&1 | tail -1
echo "== wave1"; VI_WAVE4_MIN=100000000 timeout 300 python tools/time_solver.py 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02w_pytest.log 2>&1; tail -4 gpurun_out/r02w_pytest.log
( time timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02w_bench.log 2> gpurun_out/r02w_bench.err ) 2>&1 | grep real
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02w_bench.log").read().strip().splitlines()[-1])
    print(d["metric"], d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print({k:(round(v["ms_per_step"],1)) for k,v in d["kernels"].items()})
except Exception as e: print("no line", e)
PY
( time timeout 900 python bench.py --config c3 --no-estimate > gpurun_out/r02w_c3.log 2> gpurun_out/r02w_c3.err ) 2>&1 | grep real; tail -3 gpurun_out/r02w_c3.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02w_c3.log") if l.startswith("{")][-1])
    print(d["metric"], d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print({k:(round(v["ms_per_step"],1)) for k,v in d["kernels"].items()})
    print(d.get("cpu_baseline")); print(json.dumps(d["parity"]["gpu_vs_reference"])[:600])
except Exception as e: print("no line", e)
PY
