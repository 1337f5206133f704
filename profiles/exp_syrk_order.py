"""Development aid (run under gpurun): SYRK time at the C3 / C4 shapes against the CTA order ("syrk_order": 0 k-slice major,
1 off-diagonal regions first and the cheaper diagonal regions last) and the number of CTAs per SM ("syrk_waves")."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import boom_b200  # noqa: E402

dev = torch.device("cuda:0")
shapes = ((10_000_000, 500), (1_250_000, 500), (500_000, 4000), (4_000_000, 260), (2_000_000, 1000))
for n, p in shapes:
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
    out = {"n": n, "p": p}
    for order, waves in ((0, 20), (1, 20), (1, 16), (1, 24), (1, 30), (1, 40), (0, 40)):
        ctx.set_option("syrk_order", order)
        ctx.set_option("syrk_waves", waves)
        for it in range(2):
            ctx.logit_step_device(beta, 10, 1, it, suf.data_ptr())
        ctx.synchronize(); ctx.timings(reset=True)
        reps = 3 if n * p * p > 1e12 else 6
        for it in range(reps):
            ctx.logit_step_device(beta, 10, 1, 10 + it, suf.data_ptr())
        ctx.synchronize()
        tm = ctx.timings()
        out["order%d_waves%d" % (order, waves)] = round(tm["syrk_dmma"][0] / tm["syrk_dmma"][1], 4)
    flops = float(n) * p * (p + 1) + 2.0 * n * p
    best = min(v for k, v in out.items() if k.startswith("order"))
    out["best_frac_of_37TF"] = round(flops / best * 1e-9 / 37.0, 4)
    print(json.dumps(out), flush=True)
    ctx.close()
    del X, y, aux, suf
    torch.cuda.empty_cache()
