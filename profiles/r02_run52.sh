#!/usr/bin/env bash
# Round-2 run 52: compute-sanitizer (memcheck / synccheck / initcheck) over the ragged-block SYRK kernel in both forms and the final
# Student-t kernels (racecheck omitted: ten minutes per pass; its reports on this ring are the known mbarrier-ordered ones)
set -uo pipefail
mkdir -p gpurun_out
SEL='test_ragged_diagonal_region_kernel or test_wide_p_accumulate_and_step or test_student_step_matches_oracle or test_student_loglike_matches_reference_and_oracle'
for tool in memcheck synccheck initcheck; do
  echo "=== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_student.py -m gpu -x -q -k "$SEL" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid|error" | sort | uniq -c | sort -rn | head -6
done | tee gpurun_out/r02_run52_sanitizer.log
