python -m pytest tests -m gpu -x -q > gpurun_out/r02e_pytest.log 2>&1; tail -4 gpurun_out/r02e_pytest.log
for v in "" nw6 nw4; do echo "variant [$v]"; VI_LIB_VARIANT=$v python tools/time_solver.py 28416 144 2>&1 | tail -1; done
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-estimate > gpurun_out/r02e_bench.log 2> gpurun_out/r02e_bench.err; tail -3 gpurun_out/r02e_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02e_bench.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["fit"])
print({k:(round(v["ms_per_step"],1)) for k,v in d["kernels"].items()})
PY
