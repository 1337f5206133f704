#!/usr/bin/env bash
# Round-2 run 58: active-set statistics on the Student-t spike-and-slab sampler -- the Student-t tests
set -uo pipefail
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_student.py -m gpu -x -q 2>&1 | tee gpurun_out/r02_run58_student_tests.log | tail -15
