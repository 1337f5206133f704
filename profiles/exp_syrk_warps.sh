#!/usr/bin/env bash
# Development aid (run under gpurun): the SYRK with 8 consumer warps x 2 units vs 16 x 1 unit.
set -u
for w in ${WARPS:-16}; do
  lib=/tmp/libboomgpu_cw$w.so
  make -s -C boom_b200/csrc OUT=$lib EXTRA="-DBOOMGPU_SYRK_WARPS=$w" || exit 1
  echo "== consumer warps $w"
  BOOMGPU_LIBRARY=$lib timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "accumulate or logit_step or sharding or derivatives" 2>&1 | tail -1
  BOOMGPU_LIBRARY=$lib timeout 180 python profiles/quick_perf.py c3s p128 p260 c4s 2>&1 | tail -4
done
