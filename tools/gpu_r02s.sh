timeout 300 python tools/time_estimate.py 2>&1 | tail -3
timeout 300 python tools/time_estimate.py 64 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -q -x -k "estimate or Estimate or hull or radbasfun or cli" > gpurun_out/r02s_pytest.log 2>&1; tail -3 gpurun_out/r02s_pytest.log
( time timeout 600 python bench.py --config c4 > gpurun_out/r02s_c4.log 2> gpurun_out/r02s_c4.err ) 2>&1 | grep real; tail -3 gpurun_out/r02s_c4.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02s_c4.log") if l.startswith("{")][-1])
    print(d["metric"], d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print(d["kernels"]); print({k:d["roofline"].get(k) for k in ("achieved","frac","whole_pipeline_tflops","whole_pipeline_frac")}, d["config"]["inside_hull_fraction"])
except Exception as e: print("no line", e)
PY
