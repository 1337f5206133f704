"""Experiment (run 17): the C5 kernel takes 5.7 ms driven through the C ABI in a loop and 6.35 ms inside bench.py.  Same data,
same kernel -- which difference is it?  A: C ABI loop on bench's data; B: the sampler surface, library's own stream;
C: the sampler surface on a torch stream (what bench.py does)."""
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
beta = bench.beta_true("logit", p, 5)



# A: C ABI loop
ctx = boom_b200.Context(0)
ctx.set_option("timing", 1)
ctx.set_logit_mixture(*boom_b200.default_logit_mixture())
ctx.adopt_binomial(n, p, X.data_ptr(), p, y.data_ptr(), aux.data_ptr(), keepalive=(X, y, aux))
suf = torch.empty(ctx.suf_len(), dtype=torch.float64, device=dev)
for b, tag in ((beta, "A: C ABI, beta = truth"), (np.zeros(p), "A0: C ABI, beta = 0"), (beta * 0.9, "A9: C ABI, beta = 0.9 truth")):
    for it in range(2):
        ctx.logit_step_device(b, 10, 1, it, suf.data_ptr())
    ctx.synchronize(); ctx.timings(reset=True)
    for it in range(10):
        ctx.logit_step_device(b, 10, 1, 10 + it, suf.data_ptr())
    ctx.synchronize()
    tm = ctx.timings()
    print(json.dumps({"case": tag, "kernel_ms": {k: round(v[0] / 10, 4) for k, v in tm.items() if v[1]}}), flush=True)
ctx.close()

for tag, use_stream in (("B: sampler surface, library stream", False), ("C: sampler surface, torch stream", True)):
    model = boom_b200.BinomialLogitModel(p)
    model.adopt_device_data(n, X.data_ptr(), p, y.data_ptr(), aux.data_ptr())
    prior = boom_b200.MvnModel(np.zeros(p), np.eye(p))
    s = boom_b200.BinomialLogitAuxmixSampler(model, prior, 10, boom_b200.RNG(7))
    model.set_method(s)
    if use_stream:
        st = torch.cuda.Stream(device=dev)
        model.set_stream(st.cuda_stream)
    model.set_device_option("timing", 1)
    for _ in range(3):
        model.sample_posterior()
    model.kernel_timings(True)
    for _ in range(10):
        model.sample_posterior()
    tm = model.kernel_timings(False)
    print(json.dumps({"case": tag, "timings": str(tm), "beta": [round(float(v), 3) for v in model.Beta[:6]]}), flush=True)
    del s, model
