# Round-2 evidence run (one GPU): every command first WITHOUT a profiler (its JSON line is the number that counts),
# then the ncu launch list of the bench command and one --set full capture of each solver kernel.
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
timeout 900 python bench.py --config c3 --records 16 > $O/r02_bench_c3.json 2> $O/r02_bench_c3.err; line $O/r02_bench_c3.json
timeout 600 python bench.py --config c4 > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err; line $O/r02_bench_c4.json
timeout 600 python bench.py --config c5 --records 64 > $O/r02_bench_c5.json 2> $O/r02_bench_c5.err; line $O/r02_bench_c5.json
timeout 600 python bench.py --noise-scale 1.0 --signal-terms 5 --no-cpu-baseline > $O/r02_bench_mix.json 2> $O/r02_bench_mix.err; line $O/r02_bench_mix.json
# profiler passes (numbers printed under ncu are never bench values)
python bench.py --records 2000 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > $O/r02_plain2000.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r02_launches_records2000.csv \
  python bench.py --records 2000 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > $O/r02_ncu_list.log 2>&1
python tools/time_solver.py 8192 144 > $O/r02_plain_solver.log 2>&1 || exit 1
for k in k_band k_band_tail k_chase k_tql_smem k_replay_wave; do
  ncu --set full --clock-control none --import-source on -k regex:"$k\b" -s 1 -c 1 -f -o $O/prof_r02_$k \
    python tools/time_solver.py 8192 144 > $O/r02_ncu_$k.log 2>&1
done
ls -la $O/prof_r02_* | awk '{print $5, $9}'
