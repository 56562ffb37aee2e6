# ncu of the two-stage kernels (plain run first)
python tools/time_solver.py 8192 144 > gpurun_out/r02c_plain.log 2>&1 || exit 1
tail -1 gpurun_out/r02c_plain.log
ncu --set full --clock-control none --import-source on -k regex:k_band -c 1 -f -o gpurun_out/prof_r02c_band python tools/time_solver.py 8192 144 > gpurun_out/r02c_ncu_band.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_chase -c 1 -f -o gpurun_out/prof_r02c_chase python tools/time_solver.py 8192 144 > gpurun_out/r02c_ncu_chase.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
