for o in single_launch=1 single_launch=0; do QP_OPTIONS=$o timeout 200 python profiles/quick_perf.py c1t c1h c1 c1d c1q 2>&1 | sed "s/^/$o: /" | cut -c1-220; done
