# strong scaling: 10 000 records in total over N GPUs (N from the number of visible devices)
N=$(python -c "import torch; print(torch.cuda.device_count())")
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
timeout 600 $TR bench.py --gpus $N --scaling strong --records 10000 --steps 2 --warmup 3 > gpurun_out/r02_bench_${N}gpu_strong.json 2> gpurun_out/r02_bench_${N}gpu_strong.err
python - "$N" <<'PY'
import json,sys
f=f"gpurun_out/r02_bench_{sys.argv[1]}gpu_strong.json"
d=json.loads([l for l in open(f) if l.startswith("{")][-1])
print(f, d["value"], d["unit"], "n_gpus", d["n_gpus"], d["scaling"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
