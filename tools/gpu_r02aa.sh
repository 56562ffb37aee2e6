echo "== table in global"; timeout 300 python tools/time_solver.py 2>&1 | tail -1
echo "== table in smem"; VI_WAVE_TAB_SMEM=1 timeout 300 python tools/time_solver.py 2>&1 | tail -1
echo "== small 2368"; timeout 300 python tools/time_solver.py 2368 2>&1 | tail -1
echo "== small 2368 smem"; VI_WAVE_TAB_SMEM=1 timeout 300 python tools/time_solver.py 2368 2>&1 | tail -1
echo "== n=500"; timeout 300 python tools/time_solver.py 592 500 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q -x -k "parity or tier3" > gpurun_out/r02aa_pytest.log 2>&1; tail -3 gpurun_out/r02aa_pytest.log
