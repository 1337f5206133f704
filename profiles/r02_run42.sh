#!/usr/bin/env bash
# Round-2 run 42: compute-sanitizer over what this session added -- the Student-t row model in every step kernel, the residual /
# log-likelihood kernels, the paired-diagonal SYRK form (order 2) -- then the default bench with the student_t secondary entry
set -uo pipefail
mkdir -p gpurun_out
SEL='test_student_step_matches_oracle or test_student_step_edge_cases or test_student_loglike_matches_reference_and_oracle or test_wide_p_accumulate_and_step'
for tool in memcheck racecheck synccheck initcheck; do
  echo "=== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_student.py tests/test_gpu_scale.py -m gpu -x -q -k "$SEL" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Hazard|Invalid|error" | sort | uniq -c | sort -rn | head -8
done | tee gpurun_out/r02_run42_sanitizer.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_run42_default_n1.json 2> gpurun_out/r02_run42_default_n1.err
tail -c 600 gpurun_out/r02_run42_default_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_run42_default_n1.json').read().strip().splitlines()[-1])
print('C3', d['value'], d['e2e']['value'], d['roofline']['frac'])
for k,v in d['secondary'].items(): print(k, round(v['value'],3), round(v['ms_per_step'],4), round(v['roofline']['frac'],3), v['roofline'].get('burst',{}).get('frac'))
print(d['secondary']['student_t'])
PY
