"""CPU test of the BOOM adapter's host side: with fix_latent_data(true) and the complete-data statistics pushed from
outside (the path the state-space callers of the reference use, StateSpaceLogitPosteriorSampler.cpp:58,111-123) neither
sampler imputes, so no GPU is needed; the reference's BinomialLogitSpikeSlabSampler / PoissonRegressionSpikeSlabSampler and
the B200 adapters then run only their small-state steps on the SAME statistics, on BOOM's own model objects, in one binary
(oracle/_ref/boom_adapter_demo, built here from boom_b200/boom_adapter + the compiled reference)."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "boom_adapter_demo")


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/boom_adapter_demo not built (needs the reference sources)")
@pytest.mark.parametrize("kind", ["fixed", "pfixed"])
def test_adapter_small_state_steps_match_the_reference(kind):
    iters, burn = 30000, 1000
    out = subprocess.run([EXE, kind, "400", "8", "3", str(iters), str(burn)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    ref, b2 = r["reference"], r["b200"]
    i0, i1 = np.array(ref["inclusion"]), np.array(b2["inclusion"])
    m = iters - burn
    pi = 0.5 * (i0 + i1)
    assert np.all(np.abs(i0 - i1) < 5 * np.sqrt(2 * pi * (1 - pi) * 3.0 / m) + 0.004), (i0, i1)    # sweeps on fixed statistics mix fast
    s0, s1 = np.array(ref["sd"]), np.array(b2["sd"])
    se = np.sqrt((s0 ** 2 + s1 ** 2) * 3.0 / m)
    assert np.all(np.abs(np.array(ref["mean"]) - np.array(b2["mean"])) < 5 * se + 1e-4)
    big = s0 > 0.01
    np.testing.assert_allclose(s1[big], s0[big], rtol=0.08)
