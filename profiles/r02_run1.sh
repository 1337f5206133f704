#!/usr/bin/env bash
# round 2, GPU call 1: the full -m gpu suite (incl. the new at-scale parity tests) and the default bench line
set -uo pipefail
mkdir -p gpurun_out
nproc > gpurun_out/r02_nproc.txt; free -g >> gpurun_out/r02_nproc.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r02_gputest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputest1.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02_bench1.log 2> gpurun_out/r02_bench1.err
echo "bench rc=$?" >> gpurun_out/r02_bench1.err
tail -5 gpurun_out/r02_gputest1.log; tail -c 3000 gpurun_out/r02_bench1.log; tail -5 gpurun_out/r02_bench1.err
