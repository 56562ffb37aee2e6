VI_FILL_DEBUG=2 ncu --set full --clock-control none --import-source on -k regex:"k_rows_shl_idx|k_fill_nan|k_hull_compact" -s 9 -c 3 -f -o gpurun_out/prof_r02t_rows python tools/time_estimate.py 1000 4 > gpurun_out/r02t_ncu.log 2>&1
ls -la gpurun_out/prof_r02t* | tail -2
