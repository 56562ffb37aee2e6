# Round-2 evidence run, part 2: the ncu launch list of the bench command and one --set full capture of each solver kernel,
# each after the same command has exited 0 without a profiler (numbers printed under ncu are never bench values).
set -u
O=gpurun_out
python bench.py --records 2000 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > $O/r02_plain2000.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r02_launches_records2000.csv \
  python bench.py --records 2000 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > $O/r02_ncu_list.log 2>&1
python tools/time_solver.py 8192 144 > $O/r02_plain_solver.log 2>&1 || exit 1
for k in k_band k_band_tail k_chase k_tql_smem k_replay_wave; do
  ncu --set full --clock-control none -k regex:"$k\b" -s 1 -c 1 -f -o $O/prof_r02_$k \
    python tools/time_solver.py 8192 144 > $O/r02_ncu_$k.log 2>&1
done
ls -la $O/prof_r02_* | awk '{print $5, $9}'
