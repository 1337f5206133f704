#!/usr/bin/env bash
# Round-2 run 23, 2 GPUs: the multi-GPU tests (incl. active-set under sharding) and the N = 2 bench line exactly as the driver
# launches it (final library: SYRK order / waves, active-set entries, burst re-measurement)
set -uo pipefail
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_n2_gpus.txt
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/r02_gputest23_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputest23_multi.log
tail -6 gpurun_out/r02_gputest23_multi.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02_bench23_n2.log 2> gpurun_out/r02_bench23_n2.err
echo "bench rc=$?" >> gpurun_out/r02_bench23_n2.err
tail -c 1200 gpurun_out/r02_bench23_n2.log; tail -6 gpurun_out/r02_bench23_n2.err
