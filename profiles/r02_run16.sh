#!/usr/bin/env bash
# Round-2 run 16: why is the C5 kernel 6.36 ms inside bench.py but 5.72-5.77 ms under ncu / queued back to back?
set -uo pipefail
mkdir -p gpurun_out
L=gpurun_out/r02_quick16.log; : > $L
timeout 300 python profiles/quick_perf.py c5f >> $L 2>&1
QP_MODE=sync timeout 300 python profiles/quick_perf.py c5f >> $L 2>&1
QP_MODE=sync QP_SLEEP=0.0005 timeout 300 python profiles/quick_perf.py c5f >> $L 2>&1
QP_MODE=host timeout 300 python profiles/quick_perf.py c5f >> $L 2>&1
cat $L
B="--no-e2e --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B --workload c5 --steps 20 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench c5', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])" | tee -a $L
BENCH_NO_CLOCK_SAMPLER=1 timeout 300 python bench.py $B --workload c5 --steps 20 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench c5 no clock sampler', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])" | tee -a $L
