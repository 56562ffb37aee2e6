echo "== wave1 4-deep"; timeout 300 python tools/time_solver.py 2>&1 | tail -1
echo "== small 2368"; timeout 300 python tools/time_solver.py 2368 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q -x -k "parity or tier3" > gpurun_out/r02z_pytest.log 2>&1; tail -3 gpurun_out/r02z_pytest.log
