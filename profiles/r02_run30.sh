#!/usr/bin/env bash
# Round-2 run 30: compute-sanitizer initcheck (reads of uninitialised device memory) over the parity subset + small-p variants
set -uo pipefail
mkdir -p gpurun_out
SEL='test_small_p_variants_agree_with_oracle or test_logit_step_matches_oracle or test_poisson_step_matches_oracle or test_accumulate_matches_oracle or test_loglike_derivatives_match_oracle or test_probit_step_matches_oracle or test_gather_imputer_pass_matches_the_dense_pass or test_active_set_step_matches_the_full_statistics or test_active_set_poisson_step or test_adopted_rows_that_tma_cannot_describe'
timeout 1200 compute-sanitizer --tool initcheck --print-limit 200 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q -k "$SEL" > gpurun_out/r02_initcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_initcheck.log
grep -E "Uninitialized|at .* in |Host Frame.*boomgpu_" gpurun_out/r02_initcheck.log | sed -E 's/\+0x[0-9a-f]+//; s/0x[0-9a-f]+/ADDR/g; s/thread \([0-9,]+\)/thread/; s/block \([0-9,]+\)/block/' | sort | uniq -c | sort -rn | head -30 | cut -c1-260
