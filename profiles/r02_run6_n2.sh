#!/usr/bin/env bash
# 2 GPUs: the multi-GPU tests and the N = 2 bench line exactly as the driver launches it
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_n2_gpus.txt
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/r02_gputest6_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputest6_multi.log
tail -6 gpurun_out/r02_gputest6_multi.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r02_bench6_n2.log 2> gpurun_out/r02_bench6_n2.err
echo "bench rc=$?" >> gpurun_out/r02_bench6_n2.err
tail -c 1500 gpurun_out/r02_bench6_n2.log; tail -5 gpurun_out/r02_bench6_n2.err
timeout 300 python profiles/quick_perf.py c1 c5 p32 > gpurun_out/r02_quick6.log 2>&1; cat gpurun_out/r02_quick6.log
timeout 300 python bench.py --workload c1 --steps 1000 --warmup 20 --no-secondary --no-cpu-baseline 2>/dev/null | cut -c1-330
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x -k "small_p or indicator or c5 or c2 or steps_are or logit_step or poisson_step" 2>&1 | tail -3
