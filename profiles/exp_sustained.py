"""Experiment (run 19): burst vs sustained.  (a) the copy MEASURED_PEAKS.json's hbm_gbs comes from (b.copy_(a), 1 Gi bf16), back to
back for ~1.5 s, per-launch CUDA events; (b) the C5 kernel (n = 200 M, p = 16), 160 launches back to back, per-launch events;
nvidia-smi (sm / mem clocks, power, temperature, throttle reasons) sampled every 20 ms during both."""
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import boom_b200  # noqa: E402

Q = "clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory,clocks_event_reasons.active"
rows = []
proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + Q, "--format=csv,noheader,nounits", "-lms", "20"],
                        stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)


def reader():
    for line in proc.stdout:
        rows.append((time.perf_counter(), line.strip()))


threading.Thread(target=reader, daemon=True).start()
dev = torch.device("cuda:0")


def smi_window(t0, t1):
    w = [r for t, r in rows if t0 <= t <= t1]
    return w[:: max(1, len(w) // 8)]


# (a) copy
a = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev).normal_()
b = torch.empty_like(a)
torch.cuda.synchronize()
time.sleep(1.0)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(401)]
t0 = time.perf_counter()
ev[0].record()
for i in range(400):
    b.copy_(a)
    ev[i + 1].record()
torch.cuda.synchronize()
t1 = time.perf_counter()
ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(400)])
gbs = 2 * a.numel() * 2 / ms * 1e-6
print(json.dumps({"test": "copy 2 GiB + 2 GiB", "GBps_first10": round(float(gbs[:10].mean()), 1), "GBps_by_50": [round(float(gbs[i:i + 50].mean()), 1) for i in range(0, 400, 50)],
                  "best": round(float(gbs.max()), 1), "seconds": round(t1 - t0, 2), "smi": smi_window(t0, t1)}), flush=True)
del a, b
torch.cuda.empty_cache()

# (b) the C5 kernel
n, p = 200_000_000, 16
X, y, aux = bench.make_shard("logit", n, p, 5, 0, n, dev)
torch.cuda.synchronize()
truth = bench.beta_true("logit", p, 5)
ctx = boom_b200.Context(0)
ctx.set_option("timing", 1)
ctx.set_logit_mixture(*boom_b200.default_logit_mixture())
ctx.adopt_binomial(n, p, X.data_ptr(), p, y.data_ptr(), aux.data_ptr(), keepalive=(X, y, aux))
suf = torch.empty(ctx.suf_len(), dtype=torch.float64, device=dev)
time.sleep(1.0)
times = []
t0 = time.perf_counter()
for it in range(160):
    ctx.timings(reset=True)
    ctx.logit_step_device(truth, 10, 1, it, suf.data_ptr())
    ctx.synchronize()
    times.append(ctx.timings()["fused_small"][0])
t1 = time.perf_counter()
times = np.array(times)
print(json.dumps({"test": "fused_tma_kernel C5", "ms_first5": [round(float(v), 3) for v in times[:5]], "ms_by_20": [round(float(times[i:i + 20].mean()), 3) for i in range(0, 160, 20)],
                  "seconds": round(t1 - t0, 2), "smi": smi_window(t0, t1)}), flush=True)
# after a 2 s pause
time.sleep(2.0)
times = []
for it in range(20):
    ctx.timings(reset=True)
    ctx.logit_step_device(truth, 10, 1, 200 + it, suf.data_ptr())
    ctx.synchronize()
    times.append(round(ctx.timings()["fused_small"][0], 3))
print(json.dumps({"test": "fused_tma_kernel C5 after a 2 s pause", "ms": times}), flush=True)
ctx.close()
proc.terminate()
