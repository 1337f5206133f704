#!/usr/bin/env bash
# Round-2 run 50: FINAL library -- the whole GPU suite, smoke, the default bench (both arms), launch list + DRAM traffic of the C3
# step under ncu, one --set full capture of the ragged-block kernel (p = 200)
set -uo pipefail
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_run50_smoke.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tee gpurun_out/r02_run50_gputest.log | tail -6
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_run50_default_n1.json 2> gpurun_out/r02_run50_default_n1.err
tail -c 400 gpurun_out/r02_run50_default_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_run50_reference_n1.json 2> gpurun_out/r02_run50_reference_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_run50_default_n1.json').read().strip().splitlines()[-1])
print('C3', d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('launches_per_step'), d['roofline']['kernel_ms'])
for k,v in d['secondary'].items(): print(k, round(v['value'],3), round(v['ms_per_step'],4), round(v['roofline']['frac'],3), v['roofline'].get('burst',{}).get('frac'))
print('adapter', d.get('e2e_adapter'))
r=json.loads(open('gpurun_out/r02_run50_reference_n1.json').read().strip().splitlines()[-1])
print('reference', r['value'], r['cpu_baseline'])
PY
OURS="fused_small|fused_tma|fused_ws|impute_rows|syrk_dmma|syrk_rdiag|panel_dmma|reduce_|loglike|residual|xts_|select_columns|weight_column|counts_present"
BENCH="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"$OURS" --csv --log-file gpurun_out/r02i_launches_c3.csv $BENCH > gpurun_out/r02i_launches_c3.stdout 2>&1
QP_ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:syrk_rdiag -s 1 -c 1 -f -o gpurun_out/r02i_rdiag_p200 python profiles/quick_perf.py p200 > gpurun_out/r02i_ncu_rdiag.log 2>&1
python profiles/summarize_ncu.py gpurun_out/r02i_rdiag_p200.ncu-rep > gpurun_out/r02i_rdiag_p200.summary.txt 2>/dev/null
rm -f gpurun_out/r02i_rdiag_p200.ncu-rep
ls -la gpurun_out | tail -8
