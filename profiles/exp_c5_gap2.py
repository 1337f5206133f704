"""Experiment (run 18): per-launch times of the C5 kernel, live, alternating beta = truth / 0.9 truth / 0 / perturbed -- is the
5.7 vs 6.4 ms difference a function of beta or of time?"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import boom_b200  # noqa: E402

dev = torch.device("cuda:0")
n, p = int(os.environ.get("EXP_N", 200_000_000)), 16
X, y, aux = bench.make_shard("logit", n, p, 5, 0, n, dev)
torch.cuda.synchronize()
truth = bench.beta_true("logit", p, 5)
rng = np.random.default_rng(3)
cases = [("truth", truth), ("0.9 truth", 0.9 * truth), ("truth", truth), ("zero", np.zeros(p)), ("truth + 1e-3 noise", truth + 1e-3 * rng.standard_normal(p)),
         ("truth, zeros -> 1e-4", np.where(truth == 0, 1e-4, truth)), ("0.985 truth", 0.985 * truth), ("truth", truth),
         ("2 truth", 2 * truth), ("truth * (1 + 2^-30)", truth * (1 + 2.0 ** -30))]
ctx = boom_b200.Context(0)
ctx.set_option("timing", 1)
ctx.set_logit_mixture(*boom_b200.default_logit_mixture())
ctx.adopt_binomial(n, p, X.data_ptr(), p, y.data_ptr(), aux.data_ptr(), keepalive=(X, y, aux))
suf = torch.empty(ctx.suf_len(), dtype=torch.float64, device=dev)
it = 0
for tag, b in cases:
    times = []
    for _ in range(8):
        ctx.timings(reset=True)
        ctx.logit_step_device(b, 10, 1, it, suf.data_ptr()); it += 1
        ctx.synchronize()
        times.append(round(ctx.timings()["fused_small"][0], 3))
    print(json.dumps({"beta": tag, "ms": times}), flush=True)
ctx.close()
