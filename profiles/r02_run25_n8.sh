#!/usr/bin/env bash
# 8 GPUs: the bench line exactly as the driver launches it at N = 8 (headline C3 + secondary C4, C5 + selftest)
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_n8_gpus.txt
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r02_bench25_n8.log 2> gpurun_out/r02_bench25_n8.err
echo "bench rc=$?" >> gpurun_out/r02_bench25_n8.err
tail -c 2500 gpurun_out/r02_bench25_n8.log; tail -5 gpurun_out/r02_bench25_n8.err
