#!/usr/bin/env bash
# Round-2 run 41: Student-t fast path with branch-free logarithms; the SYRK order-2 parity tests at the benchmark geometries
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_student.py -m gpu -x -q 2>&1 | tee gpurun_out/r02_run41_student_tests.log | tail -5
timeout 600 python profiles/quick_perf.py t16 t50 2>&1 | tee gpurun_out/r02_run41_student_perf.jsonl
timeout 1500 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "c3_geometry or wide_p" 2>&1 | tee gpurun_out/r02_run41_scale_tests.log | tail -5
