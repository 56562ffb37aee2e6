echo "== split"; timeout 300 python tools/time_solver.py 2>&1 | tail -1
echo "== no split"; VI_BAND_SPLIT=0 timeout 300 python tools/time_solver.py 2>&1 | tail -1
echo "== split at 5"; VI_BAND_SPLIT=5 timeout 300 python tools/time_solver.py 2>&1 | tail -1
echo "== estimate tile"; timeout 300 python tools/time_estimate.py 2>&1 | tail -2
timeout 300 python tools/time_estimate.py 64 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02m_pytest.log 2>&1; tail -4 gpurun_out/r02m_pytest.log
( time timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02m_bench.log 2> gpurun_out/r02m_bench.err ) 2>&1 | grep real
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02m_bench.log").read().strip().splitlines()[-1])
    print(d["metric"], d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    print({k:(round(v["ms_per_step"],1)) for k,v in d["kernels"].items()})
    print({k:d["roofline"].get(k) for k in ("dominant_by_time","achieved","frac","unit")})
except Exception as e: print("no line", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_hull|k_rows|k_coef|k_fill|k_est" -c 200 --csv --log-file gpurun_out/launches_r02m_est.csv python tools/time_estimate.py 1000 4 > gpurun_out/r02m_ncu_est.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_est_gemm -s 4 -c 1 -f -o gpurun_out/prof_r02m_estgemm python tools/time_estimate.py 1000 4 > gpurun_out/r02m_ncu_gemm.log 2>&1
ls -la gpurun_out/prof_r02m* | tail -2
