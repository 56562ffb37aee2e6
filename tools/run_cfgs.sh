python tools/time_solver.py 28416 144 2>&1 | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
