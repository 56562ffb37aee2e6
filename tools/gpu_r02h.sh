python -m pytest tests -m gpu -q > gpurun_out/r02h_pytest.log 2>&1; tail -5 gpurun_out/r02h_pytest.log
for cfg in "--config c1" "--config c5 --records 64 --steps 1 --warmup 1" "--config c4 --tiles 16 --e2e-tiles 4" "--config c3 --records 16 --steps 1 --warmup 1 --no-estimate --e2e-steps 1"; do
  tag=$(echo $cfg | awk '{print $2}')
  ( time timeout 600 python bench.py $cfg > gpurun_out/r02h_bench_$tag.log 2> gpurun_out/r02h_bench_$tag.err ) 2>&1 | grep real
  tail -2 gpurun_out/r02h_bench_$tag.err
  python - "$tag" <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/r02h_bench_{sys.argv[1]}.log").read().strip().splitlines()[-1])
    print(sys.argv[1], d["metric"], d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    if d.get("kernels"): print({k:(round(v["ms_per_step"],1)) for k,v in d["kernels"].items()})
    if d.get("roofline"): print({k:d["roofline"].get(k) for k in ("dominant_by_time","achieved","frac","unit")})
    if d.get("fit"): print(d["fit"])
    if d.get("cpu_baseline"): print(d["cpu_baseline"])
    if d.get("parity"): print(json.dumps(d["parity"]["gpu_vs_reference"])[:500])
except Exception as e: print("no line", e)
PY
done
