#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_draw_distributions.py tests/test_gpu_chains.py -m gpu -q -x ) > gpurun_out/r02_gputest8.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputest8.log
tail -8 gpurun_out/r02_gputest8.log
timeout 300 python profiles/quick_perf.py c2 p48 p64 > gpurun_out/r02_quick8.log 2>&1; cat gpurun_out/r02_quick8.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_ws -s 1 -c 1 -f -o gpurun_out/r02_ws_c2 python profiles/quick_perf.py c2 > gpurun_out/r02_ncu_ws.log 2>&1
python profiles/summarize_ncu.py gpurun_out/r02_ws_c2.ncu-rep > gpurun_out/r02_ws_c2.summary.txt 2>/dev/null; head -40 gpurun_out/r02_ws_c2.summary.txt
