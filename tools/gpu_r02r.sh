echo "== default"; timeout 300 python tools/time_estimate.py 2>&1 | tail -3 | head -2
echo "== no fill"; VI_FILL_DEBUG=1 timeout 300 python tools/time_estimate.py 2>&1 | tail -3 | head -2
echo "== fill on main"; VI_FILL_DEBUG=2 timeout 300 python tools/time_estimate.py 2>&1 | tail -3 | head -2
echo "== non-default main stream"; VI_TE_STREAM=1 timeout 300 python tools/time_estimate.py 2>&1 | tail -3 | head -2
