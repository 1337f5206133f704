#!/usr/bin/env bash
# Round-2 run 31: staged upload of large pageable X -- the tests that upload >= 256 MB, then the e2e leg of the headline
set -uo pipefail
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_chains.py -m gpu -x -q ) > gpurun_out/r02_gputest31.log 2>&1; tail -4 gpurun_out/r02_gputest31.log
timeout 600 python bench.py --steps 10 --warmup 5 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 value', d['value'], 'e2e', d['e2e'])"
