#!/usr/bin/env bash
# Round-2 run 39: where the SYRK's DRAM re-reads come from -- L2 promotion of the tensor map (256 / 128 / 64 B / none) and the
# off-diagonal / diagonal parts alone, time and dram__bytes_read per launch
set -uo pipefail
mkdir -p gpurun_out
CFG=1:30:3,1:30:2,1:30:1,1:30:0,2:30:0,2:60:0,1:30:3:1,1:30:3:2,1:30:0:1,1:30:0:2
EXP_CONFIGS=$CFG timeout 900 python profiles/exp_syrk_pairs.py 2>&1 | tee gpurun_out/r02_run39_syrk_promo.jsonl
EXP_REPS=1 EXP_CONFIGS=$CFG timeout 1200 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:syrk_dmma --csv --log-file gpurun_out/r02_run39_syrk_promo_ncu.csv python profiles/exp_syrk_pairs.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_run39_syrk_promo_ncu.csv')) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
out={}
for r in rows[1:]:
    out.setdefault(r[ix['ID']],{})[r[ix['Metric Name']]]=r[ix['Metric Value']]
for k,v in out.items(): print(k, v)
PY
