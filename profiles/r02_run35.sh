#!/usr/bin/env bash
# Round-2 run 35: SYRK with thread-block clusters (co-scheduled CTAs of one k-slice): time, then DRAM traffic under ncu
set -uo pipefail
mkdir -p gpurun_out
timeout 600 python profiles/exp_syrk_cluster.py 2>&1 | tee gpurun_out/r02_syrk_cluster.jsonl
EXP_REPS=1 timeout 900 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:syrk_dmma --csv --log-file gpurun_out/r02_syrk_cluster_ncu.csv python profiles/exp_syrk_cluster.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_syrk_cluster_ncu.csv')) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
out={}
for r in rows[1:]:
    out.setdefault(r[ix['ID']],{})[r[ix['Metric Name']]]=r[ix['Metric Value']]
for k,v in out.items(): print(k, v)
PY
EXP_N=500000 EXP_P=4000 EXP_CLUSTERS=0,2,4,8 EXP_REPS=2 timeout 600 python profiles/exp_syrk_cluster.py 2>&1 | tee -a gpurun_out/r02_syrk_cluster.jsonl
