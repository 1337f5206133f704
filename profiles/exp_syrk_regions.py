"""Development aid (run under gpurun): SYRK time by region class at the C3 / C4 shapes -- all regions, off-diagonal only,
diagonal only (option "syrk_filter"; results of the filtered runs are incomplete on purpose)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import boom_b200  # noqa: E402

dev = torch.device("cuda:0")
for n, p in ((2_500_000, 500), (2_500_000, 512), (500_000, 4000)):
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
    for flt, form in ((0, 0), (0, 1), (1, 0), (2, 0), (2, 1)):
        ctx.set_option("syrk_filter", flt)
        ctx.set_option("syrk_diag", form)
        for it in range(2):
            ctx.logit_step_device(beta, 10, 1, it, suf.data_ptr())
        ctx.synchronize(); ctx.timings(reset=True)
        for it in range(4):
            ctx.logit_step_device(beta, 10, 1, 10 + it, suf.data_ptr())
        ctx.synchronize()
        tm = ctx.timings()
        out["filter%d_form%d_syrk_ms" % (flt, form)] = round(tm["syrk_dmma"][0] / tm["syrk_dmma"][1], 4)
    nblk = (p + 127) // 128
    out["regions_offdiag"], out["regions_diag"] = nblk * (nblk - 1) // 2, nblk
    print(json.dumps(out), flush=True)
    ctx.close()
    del X, y, aux, suf
    torch.cuda.empty_cache()
