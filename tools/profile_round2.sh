# Round-2 evidence run, part 1 (one GPU): tests and every bench line WITHOUT a profiler (these JSON lines are the numbers
# that count).  Part 2 = tools/profile_round2_ncu.sh (gpurun returns at most 64 MiB per call).
set -u
O=gpurun_out
line() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(sys.argv[1], d.get("impl","b200"), d["value"], d["unit"], "ms/step", d.get("ms_per_step"), "e2e", d["e2e"]["value"])
except Exception as e: print(sys.argv[1], "no line", e)
PY
}
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_pytest_gpu.log 2>&1; tail -2 $O/r02_pytest_gpu.log
timeout 900 python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err || exit 1; line $O/r02_bench_default.json
timeout 900 python bench.py --impl reference > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; line $O/r02_bench_reference.json
timeout 900 python tools/parity_diag.py mid27 c1_144 c3_500 > $O/r02_parity_diag.log 2>&1; tail -3 $O/r02_parity_diag.log
timeout 600 python bench.py --config c1 > $O/r02_bench_c1.json 2> $O/r02_bench_c1.err; line $O/r02_bench_c1.json
timeout 900 python bench.py --config c3 > $O/r02_bench_c3.json 2> $O/r02_bench_c3.err; line $O/r02_bench_c3.json
timeout 600 python bench.py --config c4 > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err; line $O/r02_bench_c4.json
timeout 600 python bench.py --config c5 --records 64 > $O/r02_bench_c5.json 2> $O/r02_bench_c5.err; line $O/r02_bench_c5.json
timeout 600 python bench.py --noise-scale 1.0 --signal-terms 5 --no-cpu-baseline > $O/r02_bench_mix.json 2> $O/r02_bench_mix.err; line $O/r02_bench_mix.json
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
