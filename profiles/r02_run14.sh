#!/usr/bin/env bash
# Round-2 run 14: the full GPU suite with the v11 library, then the driver's default bench command (both arms), timed.
set -uo pipefail
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_gputest14.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputest14.log
tail -12 gpurun_out/r02_gputest14.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench14.json 2> gpurun_out/r02_bench14.err ) 2> gpurun_out/r02_bench14.time
cut -c1-600 gpurun_out/r02_bench14.json; tail -3 gpurun_out/r02_bench14.err; cat gpurun_out/r02_bench14.time
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench14_ref.json 2>> gpurun_out/r02_bench14.err ) 2>> gpurun_out/r02_bench14.time
cut -c1-600 gpurun_out/r02_bench14_ref.json; cat gpurun_out/r02_bench14.time
