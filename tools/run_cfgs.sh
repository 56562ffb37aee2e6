python -m pytest tests -m gpu -x -q > gpurun_out/pytest_packed2.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_packed2.log
for cfg in packed; do
python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_$cfg.log 2>gpurun_out/bench_$cfg.err
python - <<PY
import json
for l in open("gpurun_out/bench_$cfg.log"):
    if l.startswith("{"):
        d=json.loads(l); k=d["kernels"]
        print("$cfg:", round(d["ms_per_step"],1), "tql", round(k["tql"]["ms_per_step"],1), "apply", round(k["apply"]["ms_per_step"],1), "tridiag", round(k["tridiag"]["ms_per_step"],1), d["fit"]["status_histogram"], d["fit"]["eigen_systems_per_step"])
PY
done
