#!/usr/bin/env bash
# Round-2 run 36: the Student-t sibling on the GPU -- its tests, then the step's throughput at four shapes
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_student.py -m gpu -x -q 2>&1 | tee gpurun_out/r02_run36_student_tests.log | tail -15
timeout 600 python profiles/quick_perf.py t20 t16 t50 t500 c5 2>&1 | tee gpurun_out/r02_run36_student_perf.jsonl
