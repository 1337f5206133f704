"""Development aid (run under gpurun): the two kernels behind boomgpu_student_loglike, timed apart (CUDA events, option timing)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import boom_b200  # noqa: E402

dev = torch.device("cuda:0")
for n, p in ((25_000_000, 16), (4_000_000, 50), (1_000_000, 500)):
    g = torch.Generator(device=dev); g.manual_seed(1)
    X = torch.empty((n, p), dtype=torch.float64, device=dev).normal_(generator=g)
    y = torch.empty(n, dtype=torch.float64, device=dev).normal_(generator=g)
    beta = np.full(p, 0.01)
    ctx = boom_b200.Context(0)
    ctx.set_option("timing", 1)
    ctx.adopt_regression(n, p, X.data_ptr(), p, y.data_ptr(), keepalive=(X, y))
    ctx.student_loglike(beta, 1.5, 4.0)
    out = {}
    for what, b in (("with_beta", beta), ("stored_residuals", None)):
        ctx.synchronize(); ctx.timings(reset=True)
        t0 = time.perf_counter()
        for _ in range(10):
            ctx.student_loglike(b, 1.5, 4.0)
        wall = (time.perf_counter() - t0) / 10 * 1e3
        tm = ctx.timings()
        out[what] = {"wall_ms": round(wall, 4), "kernel_ms": {k: round(v[0] / 10, 4) for k, v in tm.items() if v[1]}}
    print(json.dumps({"n": n, "p": p, **out}), flush=True)
    ctx.close()
    del X, y
    torch.cuda.empty_cache()
