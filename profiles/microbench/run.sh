#!/usr/bin/env bash
# Run under gpurun: measures FP64 peaks on the box (results -> gpurun_out/microbench_*.jsonl)
set -euo pipefail
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_peak profiles/microbench/fp64_peak.cu
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/microbench_clocks.csv &
SMI=$!
/tmp/fp64_peak | tee gpurun_out/microbench_fp64.jsonl
python - <<'PY' | tee gpurun_out/microbench_dgemm.jsonl
import torch, json
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"test": "cublas_dgemm", "n": n, "ms": best, "tflops": 2 * n**3 / best * 1e-9}))
# weighted SYRK shape of config 3 through cuBLAS (the library baseline): (p x n) @ (n x p)
n, p = 1_000_000, 512
x = torch.randn(n, p, dtype=torch.float64, device="cuda")
torch.matmul(x.t(), x); torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(x.t(), x); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"test": "cublas_xtx_full", "n": n, "p": p, "ms": best, "tflops_full": 2 * n * p * p / best * 1e-9}))
PY
kill $SMI || true
