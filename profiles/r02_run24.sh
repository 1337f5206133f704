#!/usr/bin/env bash
# Round-2 run 24: final library -- smoke(), the full GPU suite, the driver's default bench command (both arms), timed
set -uo pipefail
mkdir -p gpurun_out
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/r02_smoke24.log 2>&1; tail -4 gpurun_out/r02_smoke24.log
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_gputest24.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputest24.log
tail -8 gpurun_out/r02_gputest24.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench24.json 2> gpurun_out/r02_bench24.err ) 2> gpurun_out/r02_bench24.time
cut -c1-300 gpurun_out/r02_bench24.json; tail -3 gpurun_out/r02_bench24.err; cat gpurun_out/r02_bench24.time
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench24_ref.json 2>> gpurun_out/r02_bench24.err ) 2>> gpurun_out/r02_bench24.time
cut -c1-200 gpurun_out/r02_bench24_ref.json; tail -4 gpurun_out/r02_bench24.time
