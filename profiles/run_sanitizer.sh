#!/usr/bin/env bash
# Run under gpurun: compute-sanitizer over a representative subset of the parity tests
# (memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards in the TMA rings; synccheck).
set -uo pipefail
mkdir -p gpurun_out
SEL='test_small_p_variants_agree_with_oracle or test_logit_step_matches_oracle or test_poisson_step_matches_oracle or test_accumulate_matches_oracle or test_loglike_derivatives_match_oracle'
for tool in memcheck racecheck synccheck; do
  echo "=== $tool"
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Hazard|Invalid|error" | tail -8
done | tee gpurun_out/sanitizer.log
