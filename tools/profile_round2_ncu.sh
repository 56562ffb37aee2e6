# Round-2 evidence run, part 2: the ncu launch list of the bench command and one --set full capture of each solver kernel,
# each after the same command has exited 0 without a profiler (numbers printed under ncu are never bench values).
set -u
O=gpurun_out
python bench.py --records 2000 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > $O/r02_plain2000.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r02_launches_records2000.csv \
  python bench.py --records 2000 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > $O/r02_ncu_list.log 2>&1
python tools/time_solver.py 8192 144 > $O/r02_plain_solver.log 2>&1 || exit 1
# (a --set full report is 14-21 MB and a call returns at most 64 MiB: the raw page of every report comes back as CSV,
#  the reports of the two tridiagonalisation kernels themselves as well)
for k in k_band k_band_tail k_chase k_tql_smem k_replay_wave; do
  ncu --set full --clock-control none -k regex:"$k\b" -s 1 -c 1 -f -o /tmp/prof_r02_$k \
    python tools/time_solver.py 8192 144 > $O/r02_ncu_$k.log 2>&1
  ncu -i /tmp/prof_r02_$k.ncu-rep --page raw --csv --print-units base > $O/prof_r02_$k.raw.csv 2>/dev/null
done
cp /tmp/prof_r02_k_band.ncu-rep /tmp/prof_r02_k_chase.ncu-rep $O/
ls -la $O/prof_r02_* | awk '{print $5, $9}'
