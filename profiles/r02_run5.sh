#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/r02_gputest5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputest5.log
tail -14 gpurun_out/r02_gputest5.log
for lib in boom_b200/libboomgpu.so tmp_libs/lib_v5B.so tmp_libs/lib_v5C.so tmp_libs/lib_v5D.so tmp_libs/lib_v5E.so; do
  echo "== $lib"
  BOOMGPU_LIBRARY=$PWD/$lib timeout 300 python profiles/quick_perf.py c5 p8 c1 p24 p32 2>&1 | tail -5
done > gpurun_out/r02_tune5.log 2>&1
cat gpurun_out/r02_tune5.log
timeout 600 python bench.py --workload c1 --steps 1000 --warmup 20 --no-secondary --no-cpu-baseline > gpurun_out/r02_bench5_c1.log 2>gpurun_out/r02_bench5_c1.err
cut -c1-400 gpurun_out/r02_bench5_c1.log
