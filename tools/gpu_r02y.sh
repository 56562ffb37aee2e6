ncu --set full --clock-control none --import-source on -k regex:"k_replay_wave4" -s 1 -c 1 -f -o gpurun_out/prof_r02y_wave4 python tools/time_solver.py 8192 144 > gpurun_out/r02y_ncu4.log 2>&1
VI_WAVE4_MIN=100000000 ncu --set full --clock-control none --import-source on -k regex:"k_replay_wave" -s 1 -c 1 -f -o gpurun_out/prof_r02y_wave1 python tools/time_solver.py 8192 144 > gpurun_out/r02y_ncu1.log 2>&1
ls -la gpurun_out/prof_r02y*
