python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest.log 2>&1; tail -6 gpurun_out/r02g_pytest.log
python bench.py > gpurun_out/r02g_bench.log 2> gpurun_out/r02g_bench.err; tail -3 gpurun_out/r02g_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02g_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["fit"])
print({k:(round(v["ms_per_step"],1)) for k,v in d["kernels"].items()})
print(json.dumps(d["roofline"])[:900])
print(json.dumps(d["parity"]["gpu_vs_reference"])[:700])
print(json.dumps(d["parity"]["gpu_vs_reference_with_gelss_driver"])[:700])
print(json.dumps(d["estimate"])[:1200])
print(d["cpu_baseline"])
PY
