#!/usr/bin/env bash
# Round-2 run 29: the driver's default bench command with the final library (parked wide tiles, e2e_adapter block), timed
set -uo pipefail
mkdir -p gpurun_out
( time timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench29.json 2> gpurun_out/r02_bench29.err ) 2> gpurun_out/r02_bench29.time
cut -c1-200 gpurun_out/r02_bench29.json; tail -3 gpurun_out/r02_bench29.err; cat gpurun_out/r02_bench29.time
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench29.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'])
print(json.dumps(d.get('e2e_adapter'))[:1500])
for k,v in d['secondary'].items(): print(k, round(v['value'],2), round(v['roofline']['frac'],4), v['roofline'].get('burst',{}).get('frac'))
PY
