# single GPU: c4 line (tiles spread over the grid), c1 line, NaN-heavy mix line, then ncu launch list + one full capture of k_est_gemm
run() { tag=$1; shift; ( time timeout 600 python bench.py "$@" > gpurun_out/r02l_$tag.log 2> gpurun_out/r02l_$tag.err ) 2>&1 | grep real; grep -v "^\[bench\|Warning\|warn" gpurun_out/r02l_$tag.err | tail -4
python - "$tag" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(f"gpurun_out/r02l_{sys.argv[1]}.log") if l.startswith("{")][-1])
    print(sys.argv[1], d["metric"], d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    if d.get("kernels"): print({k:(round(v["ms_per_step"],1)) for k,v in d["kernels"].items()})
    if d.get("roofline"): print({k:d["roofline"].get(k) for k in ("dominant_by_time","achieved","frac","unit")})
    if d.get("fit"): print(d["fit"])
    if d.get("cpu_baseline"): print(d["cpu_baseline"])
except Exception as e: print("no line", e)
PY
}
run c4 --config c4
run c1 --config c1
run mix --noise-scale 1.0 --signal-terms 5 --no-cpu-baseline
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02l_c4.csv python bench.py --config c4 --tiles 8 --e2e-tiles 2 > gpurun_out/r02l_ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_est_gemm -s 3 -c 1 -f -o gpurun_out/prof_r02l_estgemm python bench.py --config c4 --tiles 8 --e2e-tiles 2 > gpurun_out/r02l_ncu_gemm.log 2>&1
ls -la gpurun_out/prof_r02l* | tail -2
