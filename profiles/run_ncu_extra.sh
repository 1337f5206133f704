set -uo pipefail
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none -k regex:syrk_dmma -s 1 -c 1 -f -o /tmp/r01f_syrk_c4 python bench.py --workload c4 --rows 500000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r01f_ncu_syrk_c4.log 2>&1
python profiles/summarize_ncu.py /tmp/r01f_syrk_c4.ncu-rep > gpurun_out/r01f_syrk_c4.summary.txt
timeout 300 ncu --set full --clock-control none -k regex:fused_tma -s 1 -c 1 -f -o /tmp/r01f_fused_c2 python bench.py --workload c2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r01f_ncu_fused_c2.log 2>&1
python profiles/summarize_ncu.py /tmp/r01f_fused_c2.ncu-rep > gpurun_out/r01f_fused_c2.summary.txt
ls -la gpurun_out | tail -5
