( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) 2>&1 | tail -3
timeout 200 python profiles/quick_perf.py c1t c1h c1 c1d c5 c2 2>&1 | cut -c1-220
timeout 300 python bench.py --workload c1 --steps 1000 --warmup 20 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c1', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'])"
