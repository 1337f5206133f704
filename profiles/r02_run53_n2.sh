#!/usr/bin/env bash
# Round-2 run 53 (2 GPUs): the default bench line at N = 2 with the FINAL library
set -uo pipefail
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_run53_default_n2.json 2> gpurun_out/r02_run53_default_n2.err
tail -c 300 gpurun_out/r02_run53_default_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_run53_default_n2.json').read().strip().splitlines()[-1])
print('C3 N=2', d['value'], d['e2e']['value'], d['roofline']['frac'], d.get('selftest',{}).get('passed'))
for k,v in d['secondary'].items(): print(k, round(v['value'],3), round(v['ms_per_step'],4), round(v['roofline']['frac'],3))
PY
