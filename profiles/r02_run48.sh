#!/usr/bin/env bash
# Round-2 run 48: the ragged diagonal region in its own strip-form kernel -- parity, then mid-range p and C3's shape with / without it
set -uo pipefail
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "accumulate or ragged or step_matches" 2>&1 | tee gpurun_out/r02_run48_parity.log | tail -4
timeout 1200 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "wide_p or c3_geometry" 2>&1 | tee -a gpurun_out/r02_run48_parity.log | tail -4
python profiles/quick_perf.py p70 p100 p200 p260 p384 c3s 2>&1 | tee gpurun_out/r02_run48_midp.jsonl
QP_OPTIONS=syrk_rdiag=0 python profiles/quick_perf.py c3s 2>&1 | tee -a gpurun_out/r02_run48_midp.jsonl
