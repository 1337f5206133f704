#!/usr/bin/env bash
# Run under gpurun (1 GPU): the ncu evidence behind bench.py's roofline block.
#   usage: bash profiles/run_ncu.sh <round-tag>      e.g. r01
# Writes gpurun_out/<tag>_launches_{c3,c5}.csv  (every launch of this library's kernels in the bench command, with device time)
#        gpurun_out/<tag>_{syrk,impute}_c3.ncu-rep, <tag>_fused_c5.ncu-rep   (--set full captures, same workloads)
set -uo pipefail
TAG=${1:-r01}
mkdir -p gpurun_out
# only this library's kernels are profiled (torch's data-generation kernels run unprofiled, at full speed)
OURS="fused_small|fused_tma|fused_ws|impute_rows|syrk_dmma|syrk_rdiag|panel_dmma|reduce_|loglike|residual|xts_|select_columns|weight_column|counts_present"
BENCH="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary"
BENCH5="python bench.py --workload c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-secondary"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$OURS" --csv --log-file gpurun_out/${TAG}_launches_c3.csv $BENCH \
    > gpurun_out/${TAG}_launches_c3.stdout 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:syrk_dmma -s 1 -c 1 -f -o gpurun_out/${TAG}_syrk_c3 $BENCH \
    > gpurun_out/${TAG}_ncu_syrk.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:impute_rows -s 1 -c 1 -f -o gpurun_out/${TAG}_impute_c3 $BENCH \
    > gpurun_out/${TAG}_ncu_impute.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_tma -s 1 -c 1 -f -o gpurun_out/${TAG}_fused_c5 $BENCH5 \
    > gpurun_out/${TAG}_ncu_fused.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$OURS" --csv --log-file gpurun_out/${TAG}_launches_c5.csv $BENCH5 \
    > gpurun_out/${TAG}_launches_c5.stdout 2>&1
# gpurun brings back at most 64 MiB: summarise on the box, keep only the dominant kernel's report
for k in syrk_c3 impute_c3 fused_c5; do
  python profiles/summarize_ncu.py gpurun_out/${TAG}_$k.ncu-rep > gpurun_out/${TAG}_$k.summary.txt 2>/dev/null
done
rm -f gpurun_out/${TAG}_impute_c3.ncu-rep gpurun_out/${TAG}_fused_c5.ncu-rep
ls -la gpurun_out | tail -14
