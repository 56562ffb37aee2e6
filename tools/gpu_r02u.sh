echo "== n=500 two-stage (global X)"; ( time timeout 400 python tools/time_solver.py 592 500 ) 2>&1 | tail -5
echo "== n=500 one-stage"; VI_ONE_STAGE=1 timeout 400 python tools/time_solver.py 592 500 2>&1 | tail -1
echo "== n=200 two-stage (global X)"; timeout 400 python tools/time_solver.py 4096 200 2>&1 | tail -1
echo "== n=200 one-stage"; VI_ONE_STAGE=1 timeout 400 python tools/time_solver.py 4096 200 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02u_pytest.log 2>&1; tail -4 gpurun_out/r02u_pytest.log
( time timeout 600 python bench.py --config c3 --records 16 --no-estimate --no-cpu-baseline > gpurun_out/r02u_c3.log 2> gpurun_out/r02u_c3.err ) 2>&1 | grep real; tail -4 gpurun_out/r02u_c3.err
