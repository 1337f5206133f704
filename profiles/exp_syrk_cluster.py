"""Development aid (run under gpurun): the SYRK launched with thread-block clusters ("syrk_cluster" = c consecutive CTAs
co-scheduled: with the off-diagonal-first order, the regions of one k-slice that share panels) -- time by CUDA events;
run under `ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum -k regex:syrk_dmma` for the DRAM traffic per launch."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import boom_b200  # noqa: E402

dev = torch.device("cuda:0")
n, p = int(os.environ.get("EXP_N", 10_000_000)), int(os.environ.get("EXP_P", 500))
clusters = [int(c) for c in os.environ.get("EXP_CLUSTERS", "0,2,3,6").split(",")]
reps = int(os.environ.get("EXP_REPS", 4))
g = torch.Generator(device=dev); g.manual_seed(1)
X = torch.empty((n, p), dtype=torch.float64, device=dev)
step = max(1, (1 << 27) // p)
for i in range(0, n, step):
    X[i:i + step].normal_(generator=g)
y = (torch.rand(n, dtype=torch.float64, device=dev, generator=g) < 0.3).double()
aux = torch.ones(n, dtype=torch.float64, device=dev)
beta = np.zeros(p); beta[:21] = 0.1
ctx = boom_b200.Context(0)
ctx.set_option("timing", 1)
ctx.set_logit_mixture(*boom_b200.default_logit_mixture())
ctx.adopt_binomial(n, p, X.data_ptr(), p, y.data_ptr(), aux.data_ptr(), keepalive=(X, y, aux))
suf = torch.empty(ctx.suf_len(), dtype=torch.float64, device=dev)
ref = None
for c in clusters:
    ctx.set_option("syrk_cluster", c)
    ctx.logit_step_device(beta, 10, 1, 0, suf.data_ptr())
    ctx.synchronize(); ctx.timings(reset=True)
    for it in range(reps):
        ctx.logit_step_device(beta, 10, 1, 0, suf.data_ptr())
    ctx.synchronize()
    tm = ctx.timings()
    s = suf.cpu().numpy().copy()
    if ref is None:
        ref = s
    print(json.dumps({"n": n, "p": p, "cluster": c, "syrk_ms": round(tm["syrk_dmma"][0] / tm["syrk_dmma"][1], 4),
                      "same_bits_as_no_cluster": bool(np.array_equal(s, ref))}), flush=True)
ctx.close()
