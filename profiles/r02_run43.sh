#!/usr/bin/env bash
# Round-2 run 43: lane-per-row residual kernel + branch-free log likelihood for the nu draw; TRegressionSpikeSlabSampler
# (standalone + BOOM adapter); the student_t bench entry again
set -uo pipefail
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_student.py -m gpu -x -q 2>&1 | tee gpurun_out/r02_run43_student_tests.log | tail -5
timeout 1200 python -m pytest tests/test_gpu_adapter.py -m gpu -x -q -k "student" 2>&1 | tee gpurun_out/r02_run43_adapter_tests.log | tail -5
timeout 600 python profiles/bench_student.py 2>&1 | tail -1 | tee gpurun_out/r02_run43_student_bench.json
