#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x ) > gpurun_out/r02_gputest12.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputest12.log
tail -12 gpurun_out/r02_gputest12.log
timeout 300 python profiles/quick_perf.py c2 p48 p64 > gpurun_out/r02_quick12.log 2>&1; cat gpurun_out/r02_quick12.log
