#!/usr/bin/env bash
set -uo pipefail
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/r02_gputest7.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_gputest7.log
tail -25 gpurun_out/r02_gputest7.log
