#!/usr/bin/env bash
# Development aid (run under gpurun): times the small-p kernel for several (warps, ring depth) builds.
set -uo pipefail
mkdir -p gpurun_out
for cfg in ${TUNE_CFGS:-8,2,2 12,2,1 6,3,2 4,2,4}; do
  set -- ${cfg//,/ }
  lib=/tmp/libboomgpu_nw$1_s$2_r$3.so
  make -s -C boom_b200/csrc OUT=$lib EXTRA="-DBOOMGPU_TMA_NW_SMALL=$1 -DBOOMGPU_TMA_S_SMALL=$2 -DBOOMGPU_TMA_RPL_SMALL=$3" || exit 1
  echo "== NW=$1 S=$2 RPL=$3"
  BOOMGPU_LIBRARY=$lib python profiles/quick_perf.py ${TUNE_WORK:-c5 p8} 2>&1 | tail -2
done | tee gpurun_out/tune_small.log
