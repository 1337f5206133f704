#!/usr/bin/env bash
# Round-2 run 46 (2 GPUs): the multi-GPU tests incl. the Student-t sibling under sharding; the default bench line at N = 2
set -uo pipefail
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tee gpurun_out/r02_run46_gputest_multi_2gpu.log | tail -6
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_run46_default_n2.json 2> gpurun_out/r02_run46_default_n2.err
tail -c 800 gpurun_out/r02_run46_default_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_run46_default_n2.json').read().strip().splitlines()[-1])
print('C3 N=2', d['value'], d['e2e']['value'], d['roofline']['frac'], d.get('selftest'))
for k,v in d['secondary'].items(): print(k, round(v['value'],3), round(v['ms_per_step'],4), round(v['roofline']['frac'],3))
PY
