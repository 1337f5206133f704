#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_chains.py tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r02_gputest13.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputest13.log
tail -12 gpurun_out/r02_gputest13.log
B="python bench.py --no-e2e --no-cpu-baseline --no-secondary"
timeout 600 $B --steps 20 --warmup 25 --active-set > gpurun_out/r02_bench13_c3_active.log 2> gpurun_out/r02_bench13.err; cut -c1-250 gpurun_out/r02_bench13_c3_active.log
timeout 600 $B --workload c4 --steps 10 --warmup 45 --active-set > gpurun_out/r02_bench13_c4_active.log 2>> gpurun_out/r02_bench13.err; cut -c1-250 gpurun_out/r02_bench13_c4_active.log
tail -3 gpurun_out/r02_bench13.err
