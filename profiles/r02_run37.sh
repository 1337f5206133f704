#!/usr/bin/env bash
# Round-2 run 37: the Student-t sibling on the BOOM adapter (reference vs B200 sampler in one binary), the full adapter suite
# after the row_kind change, and an ncu capture of the Student-t step at p = 16
set -uo pipefail
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_adapter.py tests/test_student.py -m gpu -x -q 2>&1 | tee gpurun_out/r02_run37_adapter_tests.log | tail -15
QP_ITERS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_tma -c 1 -s 3 -o gpurun_out/r02h_student_t16 python profiles/quick_perf.py t16 > gpurun_out/r02_run37_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
