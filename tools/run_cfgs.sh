python -m pytest tests -m gpu -x -q > gpurun_out/pytest_cur.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_cur.log
python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_cur.log 2>gpurun_out/bench_cur.err
python - <<PY
import json
for l in open("gpurun_out/bench_cur.log"):
    if l.startswith("{"):
        d=json.loads(l); k=d["kernels"]
        print("step", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["ms_per_step"],1), {n: round(v["ms_per_step"],1) for n,v in k.items()}, d["fit"]["status_histogram"])
PY
