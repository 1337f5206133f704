#!/usr/bin/env bash
# Round-2 run 49: column-strip form for the off-diagonal regions of a narrow ragged last block column -- parity, then p = 200 / 260 /
# C3's and C4's shapes
set -uo pipefail
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "accumulate or ragged or step_matches" 2>&1 | tee gpurun_out/r02_run49_parity.log | tail -4
timeout 1200 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "wide_p or c3_geometry or active_set or gather" 2>&1 | tee -a gpurun_out/r02_run49_parity.log | tail -4
python profiles/quick_perf.py p200 p260 c3s c4s 2>&1 | tee gpurun_out/r02_run49_midp.jsonl
QP_OPTIONS=syrk_rdiag=0 python profiles/quick_perf.py c4s 2>&1 | tee -a gpurun_out/r02_run49_midp.jsonl
