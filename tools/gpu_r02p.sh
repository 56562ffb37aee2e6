for c in 2 4 8 16; do echo "== fill CTAs/SM $c"; VI_FILL_CTAS=$c timeout 300 python tools/time_estimate.py 2>&1 | tail -3 | head -2; done
