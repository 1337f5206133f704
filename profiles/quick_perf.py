"""Kernel-level timing of the device step on synthetic data of the BASELINE.json shapes
(development aid; bench.py is the judged harness).  Usage: python profiles/quick_perf.py [cfg ...]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import boom_b200  # noqa: E402

CFG = {
    "c1": ("logit", 100_000, 20), "c2": ("poisson", 1_000_000, 50), "c3": ("logit", 10_000_000, 500),
    "c4": ("logit", 2_000_000, 4000), "c5": ("logit", 25_000_000, 16), "c3s": ("logit", 1_000_000, 500),
    "p128": ("logit", 4_000_000, 128), "p64": ("logit", 8_000_000, 64), "p32": ("logit", 8_000_000, 32),
    "c4s": ("logit", 500_000, 4000), "p1000": ("logit", 2_000_000, 1000), "p260": ("logit", 4_000_000, 260),
    "c1t": ("logit", 9_000, 20), "c1h": ("logit", 50_000, 20), "c1d": ("logit", 200_000, 20), "c1q": ("logit", 400_000, 20),
    "p40": ("logit", 8_000_000, 40), "p48l": ("logit", 8_000_000, 48), "p56": ("logit", 8_000_000, 56), "p40p": ("poisson", 4_000_000, 40),
    "p64p": ("poisson", 4_000_000, 64), "c5m": ("logit", 100_000_000, 16), "c5f": ("logit", 200_000_000, 16), "c2x4": ("poisson", 4_000_000, 50),
    "t16": ("student", 25_000_000, 16), "t50": ("student", 4_000_000, 50), "t20": ("student", 100_000, 20), "t500": ("student", 1_000_000, 500),
    "p70": ("logit", 8_000_000, 70), "p100": ("logit", 8_000_000, 100), "p200": ("logit", 4_000_000, 200), "p384": ("logit", 2_000_000, 384),
    "p8": ("logit", 25_000_000, 8), "p24": ("logit", 12_000_000, 24), "p48": ("poisson", 4_000_000, 48),
}


def make(kind, n, p, dev):
    g = torch.Generator(device=dev); g.manual_seed(20261017)
    X = torch.empty((n, p), dtype=torch.float64, device=dev)
    step = max(1, (1 << 27) // p)
    for i in range(0, n, step):
        X[i:i + step].normal_(generator=g)
    if kind == "poisson":
        X.mul_(0.3)
    X[:, 0] = 1.0
    beta = torch.zeros(p, dtype=torch.float64, device=dev)
    beta[0] = -1.0 if kind == "logit" else 0.5
    k = min(p - 1, 20 if p >= 500 else 5)
    beta[1:1 + k] = torch.tensor([0.5 if j % 2 else -0.5 for j in range(1, 1 + k)], dtype=torch.float64, device=dev)
    eta = torch.empty(n, dtype=torch.float64, device=dev)
    for i in range(0, n, step):
        eta[i:i + step] = X[i:i + step] @ beta
    if kind == "logit":
        y = (torch.rand(n, dtype=torch.float64, device=dev, generator=g) < torch.sigmoid(eta)).double()
        aux = torch.ones(n, dtype=torch.float64, device=dev)
    elif kind == "student":   # y = eta + 1.5 t_4
        z = torch.empty(n, dtype=torch.float64, device=dev).normal_(generator=g)
        w = torch.distributions.Chi2(torch.tensor(4.0, dtype=torch.float64, device=dev)).sample((n,)) / 4.0
        y = eta + 1.5 * z / torch.sqrt(w)
        aux = None
    else:
        y = torch.poisson(torch.exp(eta), generator=g).long()
        aux = torch.ones(n, dtype=torch.float64, device=dev)
    return X, y, aux, beta.cpu().numpy()


def main():
    dev = torch.device("cuda:0")
    for name in (sys.argv[1:] or ["c1", "c2", "c3s"]):
        kind, n, p = CFG[name]
        X, y, aux, beta = make(kind, n, p, dev)
        ctx = boom_b200.Context(0)
        ctx.set_option("timing", 1)
        for kv in os.environ.get("QP_OPTIONS", "").split(","):   # e.g. QP_OPTIONS=small_variant=4,syrk_waves=24
            if "=" in kv:
                ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
        if kind == "logit":
            ctx.set_logit_mixture(*boom_b200.default_logit_mixture())
            ctx.adopt_binomial(n, p, X.data_ptr(), p, y.data_ptr(), aux.data_ptr(), keepalive=(X, y, aux))
        elif kind == "student":
            ctx.adopt_regression(n, p, X.data_ptr(), p, y.data_ptr(), keepalive=(X, y))
        else:
            ctx.set_poisson_table(*boom_b200.poisson_mixture_table_arrays())
            ctx.adopt_poisson(n, p, X.data_ptr(), p, y.data_ptr(), aux.data_ptr(), keepalive=(X, y, aux))
        suf = torch.empty(ctx.suf_len(), dtype=torch.float64, device=dev)
        iters = 3 if n * p > 1e9 else 20
        def step(it):
            if kind == "logit":
                ctx.logit_step_device(beta, 10, 1, it, suf.data_ptr())
            elif kind == "student":
                ctx.student_step_device(beta, 1.5, 4.0, 1, it, suf.data_ptr())
            else:
                ctx.poisson_step_device(beta, 1, it, suf.data_ptr())
        for it in range(2):
            step(it)
        ctx.synchronize(); ctx.timings(reset=True)
        t0 = time.perf_counter()
        mode = os.environ.get("QP_MODE", "")   # "": queued back to back; "sync": wait after every step; "host": the host-result entry
        for it in range(iters):
            if mode == "host" and kind == "logit":
                ctx.logit_step(beta, 10, 1, 10 + it)
                continue
            step(10 + it)
            if mode == "sync":
                ctx.synchronize()
                time.sleep(float(os.environ.get("QP_SLEEP", "0")))
        ctx.synchronize()
        wall = (time.perf_counter() - t0) / iters * 1e3
        tm = ctx.timings()
        per = {k: round(v[0] / max(1, iters), 4) for k, v in tm.items() if v[1]}
        bytes_ = 8.0 * n * (p + 2); flops = float(n) * p * (p + 1) + 4.0 * n * p
        dev_ms = sum(per.values())
        print(json.dumps({"mode": mode, "cfg": name, "kind": kind, "n": n, "p": p, "wall_ms": round(wall, 4), "kernel_ms": per,
                          "GBps_alg": round(bytes_ / dev_ms * 1e-6, 1), "TFLOPs_alg": round(flops / dev_ms * 1e-9, 2)}), flush=True)
        ctx.close()
        del X, y, aux, suf
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
