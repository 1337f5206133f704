#!/usr/bin/env bash
# Development aid (run under gpurun): times the SYRK for several (rows per stage, ring depth) builds.
set -uo pipefail
mkdir -p gpurun_out
for cfg in "16 6" "32 3" "8 12" "16 5"; do
  set -- $cfg
  lib=/tmp/libboomgpu_kb$1_s$2.so
  make -s -C boom_b200/csrc OUT=$lib EXTRA="-DBOOMGPU_SYRK_KB=$1 -DBOOMGPU_SYRK_STAGES=$2" || exit 1
  echo "== KB=$1 STAGES=$2"
  BOOMGPU_LIBRARY=$lib python profiles/quick_perf.py c3s p260 c4s 2>&1 | tail -3
done | tee gpurun_out/tune_syrk.log
