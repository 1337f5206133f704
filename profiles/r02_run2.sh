#!/usr/bin/env bash
# round 2, GPU call 2: parity after the SYRK trim + gather pass, C3 / C4 timings, ncu of the trimmed SYRK and the gather pass
set -uo pipefail
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/r02_gputest2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputest2.log
B="python bench.py --no-e2e --no-cpu-baseline --no-secondary"
timeout 600 $B --steps 10 --warmup 3 > gpurun_out/r02_bench2_c3.log 2> gpurun_out/r02_bench2_c3.err
timeout 600 $B --steps 10 --warmup 3 --option gather=1 > gpurun_out/r02_bench2_c3_nogather.log 2>> gpurun_out/r02_bench2_c3.err
timeout 600 $B --workload c4 --steps 4 --warmup 2 > gpurun_out/r02_bench2_c4.log 2> gpurun_out/r02_bench2_c4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:syrk_dmma -s 1 -c 1 -f -o gpurun_out/r02_syrk_c3 $B --steps 2 --warmup 1 > gpurun_out/r02_ncu_syrk.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:impute_rows -s 2 -c 1 -f -o gpurun_out/r02_impute_c3 $B --steps 3 --warmup 1 > gpurun_out/r02_ncu_impute.log 2>&1
for k in syrk_c3 impute_c3; do python profiles/summarize_ncu.py gpurun_out/r02_$k.ncu-rep > gpurun_out/r02_$k.summary.txt 2>/dev/null; done
rm -f gpurun_out/r02_impute_c3.ncu-rep
tail -3 gpurun_out/r02_gputest2.log; cat gpurun_out/r02_bench2_c3.log | cut -c1-300; ls -la gpurun_out | tail
