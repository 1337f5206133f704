#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/r02_gputest4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputest4.log
tail -12 gpurun_out/r02_gputest4.log
for lib in boom_b200/libboomgpu.so tmp_libs/lib_nw12_s2_r2.so tmp_libs/lib_nw12_s1_r2.so tmp_libs/lib_nw16_s2_r1.so tmp_libs/lib_nw16_s1_r2.so; do
  echo "== $lib"
  BOOMGPU_LIBRARY=$PWD/$lib timeout 300 python profiles/quick_perf.py c5 p8 c1 2>&1 | tail -3
done > gpurun_out/r02_tune4.log 2>&1
cat gpurun_out/r02_tune4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_tma -s 1 -c 1 -f -o gpurun_out/r02_fused_c5 python profiles/quick_perf.py c5 > gpurun_out/r02_ncu_fused.log 2>&1
python profiles/summarize_ncu.py gpurun_out/r02_fused_c5.ncu-rep > gpurun_out/r02_fused_c5.summary.txt 2>/dev/null
ls -la gpurun_out | tail -5
