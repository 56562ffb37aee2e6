python -m pytest tests -m gpu -x -q > gpurun_out/pytest_ne3.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_ne3.log
for cfg in v3; do
if [ $cfg = v2 ]; then export VI_NE_V2=1; fi
python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_$cfg.log 2>gpurun_out/bench_$cfg.err
python - <<PY
import json
for l in open("gpurun_out/bench_$cfg.log"):
    if l.startswith("{"):
        d=json.loads(l); k=d["kernels"]
        print("$cfg:", round(d["ms_per_step"],1), "ne", k["normal_eq"], "tridiag", round(k["tridiag"]["ms_per_step"],1))
PY
done
