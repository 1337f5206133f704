#!/usr/bin/env bash
# Round-2 run 40: Student-t step with the first attempts of a lane's rows interleaved -- tests, then throughput
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_student.py -m gpu -x -q 2>&1 | tee gpurun_out/r02_run40_student_tests.log | tail -8
timeout 600 python profiles/quick_perf.py t20 t16 t50 p24 2>&1 | tee gpurun_out/r02_run40_student_perf.jsonl
