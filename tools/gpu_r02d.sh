python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; tail -4 gpurun_out/r02d_pytest.log
python tools/time_solver.py 28416 144 2>&1 | tail -1
python tools/time_solver.py 600 144 2>&1 | tail -1
python tools/time_solver.py 8192 144 > gpurun_out/r02d_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_band -c 1 -f -o gpurun_out/prof_r02d_band python tools/time_solver.py 8192 144 > gpurun_out/r02d_ncu_band.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_chase -c 1 -f -o gpurun_out/prof_r02d_chase python tools/time_solver.py 8192 144 > gpurun_out/r02d_ncu_chase.log 2>&1
ls -la gpurun_out/prof_r02d*
