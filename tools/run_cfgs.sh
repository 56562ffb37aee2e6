for cfg in 8 12 24; do
VI_TQL_LANES=$cfg python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_cur.log 2>gpurun_out/bench_cur.err
python - <<PY
import json
for l in open("gpurun_out/bench_cur.log"):
    if l.startswith("{"):
        d=json.loads(l); k=d["kernels"]
        print("tql lanes $cfg: step", round(d["ms_per_step"],1), {n: round(v["ms_per_step"],1) for n,v in k.items()})
PY
done
