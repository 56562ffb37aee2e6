timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02v_pytest.log 2>&1; tail -4 gpurun_out/r02v_pytest.log
( time timeout 900 python bench.py --config c3 --no-estimate > gpurun_out/r02v_c3.log 2> gpurun_out/r02v_c3.err ) 2>&1 | grep real; tail -6 gpurun_out/r02v_c3.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02v_c3.log") if l.startswith("{")][-1])
    print(d["metric"], d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print({k:(round(v["ms_per_step"],1)) for k,v in d["kernels"].items()})
    print(d.get("cpu_baseline")); print(json.dumps(d["parity"]["gpu_vs_reference"])[:600])
except Exception as e: print("no line", e)
PY
