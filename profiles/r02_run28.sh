#!/usr/bin/env bash
# Round-2 run 28: the library with the tensor-memory-parked wide-tile form as a default -- full GPU suite, sanitizer on the
# small-p variants (memcheck + racecheck), C2 through bench.py, ncu --set full of the parked kernel at C2
set -uo pipefail
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_gputest28.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputest28.log
tail -7 gpurun_out/r02_gputest28.log
for tool in memcheck racecheck; do
  echo "=== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 50 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_p_variants or poisson_step_matches or poisson_draws" > gpurun_out/r02_san28_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r02_san28_$tool.log
  grep -E "(Write|Read) (access|Thread)|Invalid|hazard detected" gpurun_out/r02_san28_$tool.log | sed -E 's/\+0x[0-9a-f]+//; s/Thread \([0-9,]+\)/Thread/; s/0x[0-9a-f]+/ADDR/g' | sort | uniq -c | sort -rn | head -12 | cut -c1-250
done
B="--no-e2e --no-cpu-baseline --no-secondary"
timeout 300 python bench.py $B --workload c2 --steps 200 --warmup 10 2>/dev/null > gpurun_out/r02_bench28_c2.json; python -c "
import json; d=json.load(open('gpurun_out/r02_bench28_c2.json')); r=d['roofline']; print('c2', d['value'], d['ms_per_step'], r['kernel_ms'], r['frac'], r.get('fp64_frac'))"
timeout 300 ncu --set full --clock-control none -k regex:fused_tma -s 1 -c 1 -f -o /tmp/r02g_fused_c2 python bench.py $B --workload c2 --steps 2 --warmup 1 > gpurun_out/r02g_ncu_fused_c2.log 2>&1
python profiles/summarize_ncu.py /tmp/r02g_fused_c2.ncu-rep > gpurun_out/r02g_fused_c2.summary.txt; grep -E "kernel:|time_duration|pipe_tensor|issue_active|registers_per|inst_executed.sum|dram__bytes_read" gpurun_out/r02g_fused_c2.summary.txt | cut -c1-150
