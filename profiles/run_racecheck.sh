#!/usr/bin/env bash
# racecheck (default hazard analysis) with the records kept: which kernel, which source lines -- see profiles/r02_sanitizer.txt
set -uo pipefail
mkdir -p gpurun_out
SEL='test_small_p_variants_agree_with_oracle or test_logit_step_matches_oracle or test_poisson_step_matches_oracle or test_accumulate_matches_oracle or test_loglike_derivatives_match_oracle or test_probit_step_matches_oracle or test_gather_imputer_pass_matches_the_dense_pass or test_active_set_step_matches_the_full_statistics or test_active_set_poisson_step or test_adopted_rows_that_tma_cannot_describe'
timeout 900 compute-sanitizer --tool racecheck --print-limit 3000 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q -k "$SEL" > gpurun_out/racecheck_raw.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/racecheck_raw.log
# distinct (access kind, function, source line) pairs, counted
grep -E "(Write|Read) Thread" gpurun_out/racecheck_raw.log | sed -E 's/Thread \([0-9,]+\)/Thread/; s/\+0x[0-9a-f]+//' | sort | uniq -c | sort -rn | head -60 > gpurun_out/racecheck_summary.txt
grep -E "hazard detected|Race reported" gpurun_out/racecheck_raw.log | sed -E 's/0x[0-9a-f]+/ADDR/g; s/block \([0-9,]+\)/block/' | sort | uniq -c | sort -rn | head -8 >> gpurun_out/racecheck_summary.txt
cut -c1-260 gpurun_out/racecheck_summary.txt
gzip -f gpurun_out/racecheck_raw.log
