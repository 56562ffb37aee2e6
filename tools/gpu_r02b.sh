# round-2 check of the two-stage solver on the GPU box: tests, per-kernel timing A/B, parity diag, short bench
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; tail -15 gpurun_out/r02b_pytest.log
python tools/time_solver.py 28416 144 2>&1 | tail -2
VI_ONE_STAGE=1 python tools/time_solver.py 28416 144 2>&1 | tail -2
python tools/time_solver.py 600 144 2>&1 | tail -1
python tools/parity_diag.py > gpurun_out/r02b_parity_diag.log 2>&1; grep -E "gpu_vs|reference_vs" gpurun_out/r02b_parity_diag.log | cut -c1-600
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-estimate > gpurun_out/r02b_bench.log 2> gpurun_out/r02b_bench.err; tail -3 gpurun_out/r02b_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02b_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["fit"])
print({k:(round(v["ms_per_step"],1)) for k,v in d["kernels"].items()})
PY
