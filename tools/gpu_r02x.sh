echo "== wave4"; timeout 300 python tools/time_solver.py 2>&1 | tail -1
echo "== wave1"; VI_WAVE4_MIN=100000000 timeout 300 python tools/time_solver.py 2>&1 | tail -1
echo "== small 2368"; timeout 300 python tools/time_solver.py 2368 2>&1 | tail -1
echo "== n=500"; timeout 300 python tools/time_solver.py 592 500 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02x_pytest.log 2>&1; tail -3 gpurun_out/r02x_pytest.log
