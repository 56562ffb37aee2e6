echo "== wave replay"; timeout 300 python tools/time_solver.py 2>&1 | tail -2
echo "== old replay"; VI_OLD_REPLAY=1 timeout 300 python tools/time_solver.py 2>&1 | tail -2
echo "== small batch 2368"; timeout 300 python tools/time_solver.py 2368 2>&1 | tail -1
echo "== n=500, 592 systems"; ( time timeout 400 python tools/time_solver.py 592 500 ) 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/r02j_pytest.log 2>&1; tail -14 gpurun_out/r02j_pytest.log
( time timeout 600 python bench.py > gpurun_out/r02j_bench.log 2> gpurun_out/r02j_bench.err ) 2>&1 | grep real
tail -8 gpurun_out/r02j_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02j_bench.log").read().strip().splitlines()[-1])
    print(d["metric"], d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print({k:(round(v["ms_per_step"],1)) for k,v in d["kernels"].items()})
    print({k:d["roofline"].get(k) for k in ("dominant_by_time","achieved","frac","unit")})
    print(json.dumps(d["parity"]["gpu_vs_reference"])[:600])
except Exception as e: print("no line", e)
PY
( time timeout 300 python bench.py --config c3 --records 4 --steps 1 --warmup 1 --no-estimate --e2e-steps 1 --no-cpu-baseline > gpurun_out/r02j_bench_c3.log 2> gpurun_out/r02j_bench_c3.err ) 2>&1 | grep real
tail -6 gpurun_out/r02j_bench_c3.err
