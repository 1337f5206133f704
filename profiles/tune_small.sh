#!/usr/bin/env bash
# Development aid (run under gpurun): times the small-p kernel for several (warps, ring depth) builds.
set -uo pipefail
mkdir -p gpurun_out
for cfg in "12 3" "16 2" "20 2"; do
  set -- $cfg
  lib=/tmp/libboomgpu_nw$1_s$2.so
  make -s -C boom_b200/csrc OUT=$lib EXTRA="-DBOOMGPU_TMA_NW_SMALL=$1 -DBOOMGPU_TMA_S_SMALL=$2" || exit 1
  echo "== NW=$1 S=$2"
  BOOMGPU_LIBRARY=$lib python profiles/quick_perf.py c5 c1 p8 2>&1 | tail -3
done | tee gpurun_out/tune_small.log
