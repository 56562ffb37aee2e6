python -m pytest tests -m gpu -x -q > gpurun_out/pytest_sparse.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_sparse.log
for cfg in "32 32" "8 12" "4 8" "16 16" "8 6" "12 12"; do
set -- $cfg
VI_TQL_LANES=$1 VI_REPLAY_LANES=$2 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_sp_$1_$2.log 2>gpurun_out/bench_sp_$1_$2.err
python - <<PY
import json
for l in open("gpurun_out/bench_sp_$1_$2.log"):
    if l.startswith("{"):
        d=json.loads(l); k=d["kernels"]
        print("cfg $1 $2:", round(d["ms_per_step"],1), "tql", round(k["tql"]["ms_per_step"],1), "apply", round(k["apply"]["ms_per_step"],1), "tridiag", round(k["tridiag"]["ms_per_step"],1), d["fit"]["status_histogram"])
PY
done
VI_DEBUG_ROUNDS=1 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > /dev/null 2> gpurun_out/rounds.err
grep "brent round" gpurun_out/rounds.err | tail -60
