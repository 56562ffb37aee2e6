# two GPUs: default arm (weak), strong scaling, c4 point-sharded Estimate, reference arm under torchrun
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run() { tag=$1; shift; ( time timeout 600 $TR bench.py --gpus 2 "$@" > gpurun_out/r02_2gpu_$tag.log 2> gpurun_out/r02_2gpu_$tag.err ) 2>&1 | grep real; grep -v "^\[bench\|Warning\|warn" gpurun_out/r02_2gpu_$tag.err | tail -4
python - "$tag" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(f"gpurun_out/r02_2gpu_{sys.argv[1]}.log") if l.startswith("{")][-1])
    print(sys.argv[1], d.get("impl"), d["metric"], d["value"], d["unit"], "n_gpus", d["n_gpus"], d["scaling"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    if d.get("fit"): print(d["fit"])
    if d.get("estimate"): print(json.dumps(d["estimate"])[:400])
except Exception as e: print("no line", e)
PY
}
run weak
run strong --scaling strong --records 10000
run c4 --config c4 --tiles 32
run ref --impl reference --steps 1 --warmup 0
