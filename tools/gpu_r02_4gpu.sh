# four GPUs: the driver's scaling run, rehearsed (default arm, reference arm exit path, c4)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533"
( time timeout 600 $TR bench.py --gpus 4 --steps 2 --warmup 3 > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err ) 2>&1 | grep real
grep -v "^\[bench\|Warning\|warn\|^\*\|OMP_NUM" gpurun_out/r02_bench_4gpu.err | tail -5
( time timeout 600 $TR bench.py --gpus 4 --config c4 --tiles 32 > gpurun_out/r02_bench_4gpu_c4.json 2> gpurun_out/r02_bench_4gpu_c4.err ) 2>&1 | grep real
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_4gpu.json", "gpurun_out/r02_bench_4gpu_c4.json"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["unit"], "n_gpus", d["n_gpus"], d["scaling"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    except Exception as e: print(f, "no line", e)
PY
