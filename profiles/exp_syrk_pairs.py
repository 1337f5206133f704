"""Development aid (run under gpurun): the SYRK grid orders -- 1 (off-diagonal regions first, diagonal last: the round-2
default so far) against 2 (k-slice major over uniform work items: diagonal regions in pairs, super-tiled region order) --
time by CUDA events and agreement of the results; run under
`ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:syrk_dmma` for the DRAM traffic."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import boom_b200  # noqa: E402

dev = torch.device("cuda:0")
n, p = int(os.environ.get("EXP_N", 10_000_000)), int(os.environ.get("EXP_P", 500))
# order:waves[:tma_promotion[:syrk_filter]]
configs = [tuple(int(v) for v in c.split(":")) for c in os.environ.get("EXP_CONFIGS", "1:30,2:30,2:16,2:60").split(",")]
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
for cfg in configs:
    order, waves = cfg[0], cfg[1]
    promo = cfg[2] if len(cfg) > 2 else 3
    filt = cfg[3] if len(cfg) > 3 else 0
    ctx.set_option("syrk_order", order)
    ctx.set_option("syrk_waves", waves)
    ctx.set_option("tma_promotion", promo)
    ctx.set_option("syrk_filter", filt)
    ctx.logit_step_device(beta, 10, 1, 0, suf.data_ptr())
    ctx.synchronize(); ctx.timings(reset=True)
    for it in range(reps):
        ctx.logit_step_device(beta, 10, 1, 0, suf.data_ptr())
    ctx.synchronize()
    tm = ctx.timings()
    s = suf.cpu().numpy().copy()
    if ref is None:
        ref = s
    d = np.sqrt(np.abs(np.diag(ref[:p * p].reshape(p, p))))
    err = float(np.max(np.abs(s[:p * p] - ref[:p * p]).reshape(p, p) / np.outer(d, d)))
    print(json.dumps({"n": n, "p": p, "order": order, "waves": waves, "tma_promotion": promo, "filter": filt, "syrk_ms": round(tm["syrk_dmma"][0] / reps, 4),   # per step: main grid + ragged-block kernel
                      "syrk_launches_per_step": tm["syrk_dmma"][1] / reps,
                      "reduce_ms": round(tm["reduce"][0] / reps, 4),
                      "normwise_diff_vs_first": err, "xty_diff": float(np.max(np.abs(s[p * p:] - ref[p * p:])))}), flush=True)
ctx.close()
