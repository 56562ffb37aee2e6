timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest.log 2>&1; tail -2 gpurun_out/r02_final_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --config c4 > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err; tail -2 gpurun_out/r02_bench_c4.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_c4.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["steps"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["kernel_alone"]["frac"], d["config"]["full_grid_timed"], d["clocks"])
PY
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -2 gpurun_out/r02_bench_default.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_default.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["traffic"], d["cpu_baseline"]["value"], d["gpu_launches"])
PY
