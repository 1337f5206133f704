#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py tests/test_draw_distributions.py -m gpu -q ) > gpurun_out/r02_gputest3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputest3.log
timeout 600 python profiles/exp_syrk_regions.py > gpurun_out/r02_exp_syrk_regions.log 2>&1
timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-secondary --steps 10 --warmup 3 > gpurun_out/r02_bench3_c3.log 2> gpurun_out/r02_bench3_c3.err
tail -3 gpurun_out/r02_gputest3.log; cat gpurun_out/r02_exp_syrk_regions.log; cut -c1-200 gpurun_out/r02_bench3_c3.log
timeout 600 python profiles/quick_perf.py c5 p8 c1 c2 p32 p64 > gpurun_out/r02_quick3.log 2>&1
cat gpurun_out/r02_quick3.log
