#!/usr/bin/env bash
# Development aid (run under gpurun): the single-pass kernel at 40 < p <= 64 for several (warps, ring depth) builds.
set -uo pipefail
mkdir -p gpurun_out
for cfg in ${TUNE_CFGS:-6,2 8,1 12,1 10,1}; do
  set -- ${cfg//,/ }
  lib=/tmp/libboomgpu_w$1_s$2.so
  make -s -C boom_b200/csrc OUT=$lib EXTRA="-DBOOMGPU_TMA_NW_WIDE=$1 -DBOOMGPU_TMA_S_WIDE=$2" || exit 1
  echo "== NW=$1 S=$2"
  BOOMGPU_LIBRARY=$lib python profiles/quick_perf.py ${TUNE_WORK:-p48 p64 c2} 2>&1 | tail -3
done | tee gpurun_out/tune_wide.log
