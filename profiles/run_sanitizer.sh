#!/usr/bin/env bash
# Run under gpurun: compute-sanitizer over a representative subset of the parity tests
# (memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards in the TMA rings; synccheck).
# Round 2 adds the kernels written this round: the single-launch tail, the gather imputer pass, the strip / balanced SYRK forms
# (p = 500 and p = 520 shapes), the active-set panel product and column fetch, the probit step, the warp-specialised variant.
set -uo pipefail
mkdir -p gpurun_out
SEL='test_small_p_variants_agree_with_oracle or test_logit_step_matches_oracle or test_poisson_step_matches_oracle or test_accumulate_matches_oracle or test_loglike_derivatives_match_oracle or test_probit_step_matches_oracle or test_gather_imputer_pass_matches_the_dense_pass or test_active_set_step_matches_the_full_statistics or test_active_set_poisson_step or test_adopted_rows_that_tma_cannot_describe'
for tool in memcheck racecheck synccheck; do
  echo "=== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q -k "$SEL" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Hazard|Invalid|error" | sort | uniq -c | sort -rn | head -12
done | tee gpurun_out/sanitizer.log
