for cfg in "8 16" "12 16" "8 24" "6 32" "12 12"; do
set -- $cfg
VI_REPLAY_WARPS=$1 VI_REPLAY_LANES=$2 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_cur.log 2>gpurun_out/bench_cur.err
python - <<PY
import json
for l in open("gpurun_out/bench_cur.log"):
    if l.startswith("{"):
        d=json.loads(l); k=d["kernels"]
        print("cfg $1 $2: step", round(d["ms_per_step"],1), {n: round(v["ms_per_step"],1) for n,v in k.items()})
PY
done
