#!/usr/bin/env bash
# Round-2 run 27: tensor-memory-parked accumulators, after the register diet (small_variant 4: 12 warps) -- parity, then time
set -uo pipefail
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_p_variants" ) > gpurun_out/r02_gputest27.log 2>&1; tail -3 gpurun_out/r02_gputest27.log
L=gpurun_out/r02_quick27.log; : > $L
for v in 0 4; do
  QP_OPTIONS=small_variant=$v timeout 400 python profiles/quick_perf.py c2 c2x4 p40p p48 p64p p40 p48l p56 p64 2>&1 | sed "s/^/variant $v: /" >> $L
done
python - <<'PY'
import json
rows={}
for line in open('gpurun_out/r02_quick27.log'):
    if not line.startswith('variant'): continue
    v=int(line.split(':')[0].split()[1]); d=json.loads(line.split(': ',1)[1])
    rows.setdefault(d['cfg'],{})[v]=d['kernel_ms']['fused_small']
for k,r in rows.items(): print(k, r, 'ratio %.3f' % (r.get(4,0)/r.get(0,1)))
PY
