#!/usr/bin/env bash
# Round-2 run 26: tensor-memory-parked accumulators for wide tiles (small_variant 4: 12 warps, 5: 10 warps) -- parity, then time
set -uo pipefail
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_p_variants" ) > gpurun_out/r02_gputest26.log 2>&1; tail -5 gpurun_out/r02_gputest26.log
L=gpurun_out/r02_quick26.log; : > $L
for v in 0 4 5; do
  QP_OPTIONS=small_variant=$v timeout 300 python profiles/quick_perf.py c2 c2x4 p48 p64 2>&1 | sed "s/^/variant $v: /" >> $L
done
cat $L
