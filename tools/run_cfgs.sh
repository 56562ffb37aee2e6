for b in 1 2 3 4 6; do echo -n "batch $b: "; VI_TQL_BATCH=$b python tools/time_solver.py 28416 144 2>&1 | tail -1; done
