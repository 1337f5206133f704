#!/usr/bin/env bash
# Round-2 run 15: final ncu evidence (launch lists + --set full captures of the dominant kernels) and a row-count sweep of the
# small-p kernel (is the per-row cost at 200 M rows the one measured at 25 M?).
set -uo pipefail
mkdir -p gpurun_out
timeout 600 python profiles/quick_perf.py c5 c5m c5f c2 c2x4 c1 > gpurun_out/r02_quick15.log 2>&1; cat gpurun_out/r02_quick15.log
bash profiles/run_ncu.sh r02f > gpurun_out/r02f_run_ncu.log 2>&1; tail -8 gpurun_out/r02f_run_ncu.log
B="--no-e2e --no-cpu-baseline --no-secondary"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:panel_dmma -s 22 -c 1 -f -o /tmp/r02f_panel_c3 python bench.py $B --active-set --steps 2 --warmup 25 > gpurun_out/r02f_ncu_panel.log 2>&1
python profiles/summarize_ncu.py /tmp/r02f_panel_c3.ncu-rep > gpurun_out/r02f_panel_c3.summary.txt
timeout 300 ncu --set full --clock-control none -k regex:fused_tma -s 1 -c 1 -f -o /tmp/r02f_fused_c2 python bench.py $B --workload c2 --steps 2 --warmup 1 > gpurun_out/r02f_ncu_fused_c2.log 2>&1
python profiles/summarize_ncu.py /tmp/r02f_fused_c2.ncu-rep > gpurun_out/r02f_fused_c2.summary.txt
timeout 400 ncu --set full --clock-control none -k regex:syrk_dmma -s 1 -c 1 -f -o /tmp/r02f_syrk_c4 python bench.py $B --workload c4 --rows 500000 --steps 2 --warmup 1 > gpurun_out/r02f_ncu_syrk_c4.log 2>&1
python profiles/summarize_ncu.py /tmp/r02f_syrk_c4.ncu-rep > gpurun_out/r02f_syrk_c4.summary.txt
ls -la gpurun_out | tail -16
