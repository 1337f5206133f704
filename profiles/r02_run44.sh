#!/usr/bin/env bash
# Round-2 run 44: residual kernel with G lanes per row -- Student-t tests, the student_t bench entry
set -uo pipefail
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_student.py -m gpu -x -q 2>&1 | tee gpurun_out/r02_run44_student_tests.log | tail -5
timeout 600 python profiles/bench_student.py 2>&1 | tail -1 | tee gpurun_out/r02_run44_student_bench.json
