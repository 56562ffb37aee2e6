# Round profile: plain run first (must exit 0), then the ncu launch list, then --set full captures of the top kernels
# (one launch each: gpurun returns at most 64 MiB).
python bench.py --records 300 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/plain_r01b.log 2> gpurun_out/plain_r01b.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r01b.csv \
  python bench.py --records 300 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
for k in k_tridiag_packed k_tql_smem k_replay k_ne_dmma3; do
  ncu --set full --clock-control none -k regex:$k -s 2 -c 1 -o gpurun_out/prof_r01b_$k -f \
    python bench.py --records 2000 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out; echo done
