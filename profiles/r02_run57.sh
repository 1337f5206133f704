#!/usr/bin/env bash
# Round-2 run 57: the library as committed at the end of the round -- smoke, the whole GPU suite, the default bench, the reference arm
set -uo pipefail
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_run57_smoke.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tee gpurun_out/r02_run57_gputest.log | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_run57_default_n1.json 2> gpurun_out/r02_run57_default_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_run57_reference_n1.json 2> gpurun_out/r02_run57_reference_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_run57_default_n1.json').read().strip().splitlines()[-1])
print('C3', d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'])
for k,v in d['secondary'].items(): print(k, round(v['value'],3), round(v['ms_per_step'],4), round(v['roofline']['frac'],3))
r=json.loads(open('gpurun_out/r02_run57_reference_n1.json').read().strip().splitlines()[-1])
print('reference', r['value'])
PY
