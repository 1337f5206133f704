#!/usr/bin/env bash
# Round-2 run 38: SYRK grid order 2 (uniform work items: diagonal regions in pairs) against order 1 -- time, then DRAM traffic
set -uo pipefail
mkdir -p gpurun_out
timeout 900 python profiles/exp_syrk_pairs.py 2>&1 | tee gpurun_out/r02_run38_syrk_pairs.jsonl
EXP_REPS=1 EXP_CONFIGS=1:30,2:30,2:16 timeout 900 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:syrk_dmma --csv --log-file gpurun_out/r02_run38_syrk_pairs_ncu.csv python profiles/exp_syrk_pairs.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_run38_syrk_pairs_ncu.csv')) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
out={}
for r in rows[1:]:
    out.setdefault(r[ix['ID']],{})[r[ix['Metric Name']]]=r[ix['Metric Value']]
for k,v in out.items(): print(k, v)
PY
EXP_N=500000 EXP_P=4000 EXP_REPS=2 EXP_CONFIGS=1:30,2:30,2:8 timeout 600 python profiles/exp_syrk_pairs.py 2>&1 | tee -a gpurun_out/r02_run38_syrk_pairs.jsonl
EXP_N=500000 EXP_P=4000 EXP_REPS=1 EXP_CONFIGS=1:30,2:30 timeout 900 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:syrk_dmma --csv --log-file gpurun_out/r02_run38_syrk_pairs_c4_ncu.csv python profiles/exp_syrk_pairs.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_run38_syrk_pairs_c4_ncu.csv')) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
out={}
for r in rows[1:]:
    out.setdefault(r[ix['ID']],{})[r[ix['Metric Name']]]=r[ix['Metric Value']]
for k,v in out.items(): print(k, v)
PY
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q -k "accumulate or syrk or step" 2>&1 | tail -5
