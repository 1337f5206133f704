#!/usr/bin/env bash
# Round-2 run 22: SYRK CTA order / waves experiment, the new tests (Poisson active-set, order invariance), racecheck records
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python profiles/exp_syrk_order.py > gpurun_out/r02_syrk_order.jsonl 2>&1; cat gpurun_out/r02_syrk_order.jsonl | cut -c1-400
( timeout 900 python -m pytest tests/test_gpu_chains.py tests/test_gpu_adapter.py tests/test_gpu_scale.py -m gpu -x -q -k "active or wide_p or c3_geometry" ) > gpurun_out/r02_gputest22.log 2>&1; tail -5 gpurun_out/r02_gputest22.log
bash profiles/run_racecheck.sh
