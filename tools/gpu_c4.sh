timeout 600 python bench.py --config c4 > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err; tail -2 gpurun_out/r02_bench_c4.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_c4.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["kernel_alone"])
PY
