"""Two-GPU test of the sharded path (skipped on a one-GPU box): one process per GPU, NCCL all-reduce of the
packed statistics through boom_b200.distributed, against the single-GPU run on the same data.
  * the all-reduced statistics equal the one-GPU statistics (same draws: Philox keyed by the global row)
  * the chains of the two ranks are identical, and equal to the one-GPU chain up to summation order."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _chain(model_factory, p, iters, attach=None):
    import boom_b200
    model = model_factory()
    sampler = boom_b200.BinomialLogitSpikeSlabSampler(model, boom_b200.MvnModel(np.zeros(p), np.eye(p)),
                                                      boom_b200.VariableSelectionPrior(p, 0.3), 10, boom_b200.RNG(17))
    model.set_method(sampler)
    if attach:
        attach(model)
    out, sufs = [], []
    for _ in range(iters):
        model.sample_posterior()
        out.append(np.array(model.Beta))
        sufs.append(np.array(sampler.suf.xtx))
    return np.array(out), np.array(sufs), sampler.suf.sample_size


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import boom_b200
    from boom_b200 import distributed as shard
    from oracle import oracle as O
    try:
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=dev)
        n, p = 40_001, 70          # p > 64: the two-pass TMA/DMMA path; odd n: uneven shards
        X, y, nt, _ = O.synth_binomial(n, p, 4, seed=3, max_trials=2)
        row0, row1 = shard.shard_range(n, world, rank)
        stream = torch.cuda.Stream(device=dev)
        hooks = []

        def attach(model):
            hooks.append(shard.attach(model, n, stream, dev, rank, world)[2])
        betas, sufs, ss = _chain(lambda: boom_b200.BinomialLogitModel(X[row0:row1], y[row0:row1], nt[row0:row1]), p, 12, attach)
        assert ss == n and hooks[0].calls == 12
        q.put((rank, betas, sufs))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        q.put((rank, "FAIL: %r" % (e,), None))


def test_two_gpu_chain_matches_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    import boom_b200
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
    assert not isinstance(res[0][1], str) and not isinstance(res[1][1], str), res
    np.testing.assert_array_equal(res[0][1], res[1][1])      # identical chains on both ranks
    np.testing.assert_array_equal(res[0][2], res[1][2])      # identical all-reduced statistics
    n, p = 40_001, 70
    X, y, nt, _ = O.synth_binomial(n, p, 4, seed=3, max_trials=2)
    betas1, sufs1, ss1 = _chain(lambda: boom_b200.BinomialLogitModel(X, y, nt), p, 12)
    # first iteration: same beta (zeros) -> same draws -> statistics equal up to summation order
    d = np.sqrt(np.diag(sufs1[0]))
    assert np.max(np.abs(res[0][2][0] - sufs1[0]) / np.outer(d, d)) < 1e-12
    # the whole chain stays together (the host steps see statistics that differ in the last bits only)
    np.testing.assert_allclose(res[0][1], betas1, rtol=1e-6, atol=1e-8)


def _native_worker(rank, world, uid, q):
    sys.path.insert(0, ROOT)
    import boom_b200
    from boom_b200.distributed import shard_range
    from oracle import oracle as O
    try:
        n, p = 30_011, 24
        X, y, nt, beta = O.synth_binomial(n, p, 4, seed=9, max_trials=1)
        row0, row1 = shard_range(n, world, rank)
        ctx = boom_b200.Context(rank)
        mix = O.logit_mixture()
        ctx.set_logit_mixture(mix.mu, mix.sigma, mix.weights)
        ctx.upload_binomial(X[row0:row1], y[row0:row1], nt[row0:row1])
        ctx.set_row_offset(row0)
        ctx.comm_init(uid, world, rank)            # ncclCommInitRank inside libboomgpu (NCCL bound at run time)
        xtx, xty, ss = ctx.logit_step(beta, 10, seed=21, iteration=4)   # all-reduces before the copy to the host
        ctx.comm_destroy()
        ctx.close()
        q.put((rank, xtx, xty, ss))
    except Exception as e:  # noqa: BLE001
        q.put((rank, "FAIL: %r" % (e,), None, None))


def test_native_nccl_allreduce_through_the_c_abi():
    """boomgpu_comm_unique_id / boomgpu_comm_init / the all-reducing boomgpu_logit_step: no torch in the loop."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    import boom_b200
    from oracle import oracle as O
    uid = boom_b200.Context.comm_unique_id()
    assert len(uid) == 128
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_native_worker, args=(r, 2, uid, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
    assert not isinstance(res[0][1], str) and not isinstance(res[1][1], str), res
    np.testing.assert_array_equal(res[0][1], res[1][1])
    n, p = 30_011, 24
    X, y, nt, beta = O.synth_binomial(n, p, 4, seed=9, max_trials=1)
    rxtx, rxty, rss, _ = O.logit_step(X, y, nt, beta, 10, O.logit_mixture(), 21, 4)
    d = np.sqrt(np.diag(rxtx))
    assert res[0][3] == rss == n
    assert np.max(np.abs(res[0][1] - rxtx) / np.outer(d, d)) < 1e-11
    assert np.max(np.abs(res[0][2] - rxty)) < 1e-9 * np.max(np.abs(rxty))


def _native_attach_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import boom_b200
    from boom_b200 import distributed as shard
    from oracle import oracle as O
    try:
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=dev)
        n, p = 20_003, 10
        X, y, nt, _ = O.synth_binomial(n, p, 3, seed=4)
        row0, row1 = shard.shard_range(n, world, rank)
        stream = torch.cuda.Stream(device=dev)
        betas, sufs, ss = _chain(lambda: boom_b200.BinomialLogitModel(X[row0:row1], y[row0:row1], nt[row0:row1]), p, 8,
                                 lambda m: shard.attach(m, n, stream, dev, rank, world, native=True))
        assert ss == n
        q.put((rank, betas, sufs))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        q.put((rank, "FAIL: %r" % (e,), None))


def test_sampler_surface_with_the_native_communicator():
    """distributed.attach(native=True): the id travels through torch.distributed, the all-reduce runs inside the C ABI step."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_native_attach_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
    assert not isinstance(res[0][1], str) and not isinstance(res[1][1], str), res
    np.testing.assert_array_equal(res[0][1], res[1][1])
    np.testing.assert_array_equal(res[0][2], res[1][2])


def _loglike_worker(rank, world, port, native, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import boom_b200
    from boom_b200 import distributed as shard
    from oracle import oracle as O
    try:
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=dev)
        n, p = 20_003, 9
        X, y, nt, beta = O.synth_binomial(n, p, 3, seed=14, max_trials=3)
        row0, row1 = shard.shard_range(n, world, rank)
        stream = torch.cuda.Stream(device=dev)
        model = boom_b200.BinomialLogitModel(X[row0:row1], y[row0:row1], nt[row0:row1])
        s = boom_b200.BinomialLogitSpikeSlabSampler(model, boom_b200.MvnModel(np.zeros(p), np.eye(p)),
                                                    boom_b200.VariableSelectionPrior(p, 0.5), 10, boom_b200.RNG(3))
        model.set_method(s)
        shard.attach(model, n, stream, dev, rank, world, native=native)
        ll = model.log_likelihood(beta * 0.8)
        ll2, g, h = model.log_likelihood_derivs(beta * 0.8)
        model.drop_all()
        for j in range(4):
            model.add(j)
        s.find_posterior_mode(1e-8)
        q.put((rank, ll, ll2, g, h, np.array(model.Beta), s.log_posterior_at_mode))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        q.put((rank, "FAIL: %r" % (e,), None, None, None, None, None))


@pytest.mark.parametrize("native", [False, True])
def test_log_likelihood_and_mode_are_all_reduced_under_sharding(native):
    """with rows sharded, log_likelihood / log_likelihood_derivs / find_posterior_mode must see ALL rows on every
    rank (hook path and native communicator): equal on both ranks and equal to the one-GPU values."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    import boom_b200
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_loglike_worker, args=(r, 2, port, native, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
    assert not isinstance(res[0][1], str) and not isinstance(res[1][1], str), res
    for k in range(1, 7):
        np.testing.assert_array_equal(res[0][k], res[1][k])
    n, p = 20_003, 9
    X, y, nt, beta = O.synth_binomial(n, p, 3, seed=14, max_trials=3)
    rll, rg, rh = O.binomial_logit_loglike_derivs(X, y, nt, beta * 0.8)
    assert res[0][1] == pytest.approx(rll, rel=1e-12) and res[0][2] == pytest.approx(rll, rel=1e-12)
    np.testing.assert_allclose(res[0][3], rg, rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(res[0][4], rh, rtol=1e-10, atol=1e-9)
    model = boom_b200.BinomialLogitModel(X, y, nt)
    s = boom_b200.BinomialLogitSpikeSlabSampler(model, boom_b200.MvnModel(np.zeros(p), np.eye(p)),
                                                boom_b200.VariableSelectionPrior(p, 0.5), 10, boom_b200.RNG(3))
    model.set_method(s)
    model.drop_all()
    for j in range(4):
        model.add(j)
    s.find_posterior_mode(1e-8)
    np.testing.assert_allclose(res[0][5], np.array(model.Beta), rtol=1e-7, atol=1e-9)
    assert res[0][6] == pytest.approx(s.log_posterior_at_mode, rel=1e-10)


def _active_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import boom_b200
    from boom_b200 import distributed as shard
    from oracle import oracle as O
    try:
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=dev)
        n, p = 30_001, 150
        X, y, nt, _ = O.synth_binomial(n, p, 6, seed=77, max_trials=1)
        row0, row1 = shard.shard_range(n, world, rank)
        stream = torch.cuda.Stream(device=dev)
        chains, fetched = [], 0
        for active in (False, True):
            model = boom_b200.BinomialLogitModel(X[row0:row1], y[row0:row1], nt[row0:row1])
            s = boom_b200.BinomialLogitSpikeSlabSampler(model, boom_b200.MvnModel(np.zeros(p), np.eye(p)),
                                                        boom_b200.VariableSelectionPrior(p, 6.0 / p), 10, boom_b200.RNG(21))
            s.set_active_set_statistics(active)
            model.set_method(s)
            shard.attach(model, n, stream, dev, rank, world, native=True)
            model.drop_all()
            model.add(0)
            out = []
            for _ in range(30):
                model.sample_posterior()
                out.append(np.array(model.Beta))
            chains.append(np.array(out))
            if active:
                fetched = s.active_set_columns_fetched
                assert s.suf.sample_size == n
            del s, model
        q.put((rank, chains[0], chains[1], fetched))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        q.put((rank, "FAIL: %r" % (e,), None, 0))


def test_active_set_statistics_under_sharding():
    """The active-set option with the native communicator: the panel product, the fetched columns and the on-demand full
    statistics are all-reduced inside the C ABI; both ranks walk the same chain, which is the full-statistics chain."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_active_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
    assert not isinstance(res[0][1], str) and not isinstance(res[1][1], str), res
    np.testing.assert_array_equal(res[0][2], res[1][2])                      # the two ranks: identical
    assert np.array_equal(res[0][1] != 0, res[0][2] != 0)                    # same decisions as the full-statistics chain
    np.testing.assert_allclose(res[0][2], res[0][1], rtol=1e-7, atol=1e-9)
    assert res[0][3] >= 6


def _student_worker(rank, world, uid, q):
    sys.path.insert(0, ROOT)
    import boom_b200
    from boom_b200.distributed import shard_range
    from oracle import oracle as O
    try:
        n, p = 30_011, 9
        X, y, bt = O.synth_student(n, p, 3, 41)
        row0, row1 = shard_range(n, world, rank)
        # (1) the C ABI: all-reduced statistics and log likelihood (with beta, then from the stored residuals)
        ctx = boom_b200.Context(rank)
        ctx.upload_regression(X[row0:row1], y[row0:row1])
        ctx.set_row_offset(row0)
        ctx.comm_init(uid[0], world, rank)
        xtwx, xtwy, sc = ctx.student_step(bt * 0.9, 1.4, 3.5, 31, 2)
        ll1 = ctx.student_loglike(bt * 0.9, 1.4, 3.5)
        ll2 = ctx.student_loglike(None, 1.1, 8.0)
        ctx.comm_destroy()
        ctx.close()
        # (2) the sampler surface: the same chain on both ranks
        model = boom_b200.TRegressionModel(X[row0:row1], y[row0:row1])
        s = boom_b200.TRegressionSpikeSlabSampler(model, boom_b200.MvnModel(np.zeros(p), 4.0 * np.eye(p)),
                                                  boom_b200.VariableSelectionPrior(p, 0.4), boom_b200.ChisqModel(1.0, 1.0),
                                                  boom_b200.UniformModel(0.5, 60.0), boom_b200.RNG(19))
        model.set_method(s)
        model.set_device(rank)
        model.set_row_offset(row0)
        model.set_communicator(uid[1], world, rank)
        chain = []
        for _ in range(15):
            model.sample_posterior()
            chain.append(np.r_[model.Beta, model.sigsq, model.nu])
        q.put((rank, xtwx, xtwy, sc, ll1, ll2, np.array(chain)))
    except Exception as e:  # noqa: BLE001
        q.put((rank, "FAIL: %r" % (e,), None, None, None, None, None))


def test_student_t_sibling_under_sharding():
    """The Student-t step and its log likelihood all-reduce natively (statistics and likelihood of ALL rows on every rank, equal
    to one context holding all rows), and the sharded TRegressionSpikeSlabSampler chain is identical on both ranks and equal to
    the one-GPU chain up to summation order."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    import boom_b200
    from oracle import oracle as O
    uid = [boom_b200.Context.comm_unique_id(), boom_b200.Context.comm_unique_id()]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_student_worker, args=(r, 2, uid, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
    assert not isinstance(res[0][1], str) and not isinstance(res[1][1], str), res
    for k in range(1, 7):
        np.testing.assert_array_equal(res[0][k], res[1][k])
    n, p = 30_011, 9
    X, y, bt = O.synth_student(n, p, 3, 41)
    one = boom_b200.Context(0)
    one.upload_regression(X, y)
    xtwx, xtwy, sc = one.student_step(bt * 0.9, 1.4, 3.5, 31, 2)
    d = np.sqrt(np.diag(xtwx))
    assert np.max(np.abs(res[0][1] - xtwx) / np.outer(d, d)) < 1e-12
    np.testing.assert_allclose(res[0][2], xtwy, rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(res[0][3], sc, rtol=1e-11)
    assert res[0][4] == pytest.approx(one.student_loglike(bt * 0.9, 1.4, 3.5), rel=1e-12)
    assert res[0][5] == pytest.approx(one.student_loglike(None, 1.1, 8.0), rel=1e-12)
    one.close()
    model = boom_b200.TRegressionModel(X, y)
    s = boom_b200.TRegressionSpikeSlabSampler(model, boom_b200.MvnModel(np.zeros(p), 4.0 * np.eye(p)),
                                              boom_b200.VariableSelectionPrior(p, 0.4), boom_b200.ChisqModel(1.0, 1.0),
                                              boom_b200.UniformModel(0.5, 60.0), boom_b200.RNG(19))
    model.set_method(s)
    chain = []
    for _ in range(15):
        model.sample_posterior()
        chain.append(np.r_[model.Beta, model.sigsq, model.nu])
    np.testing.assert_allclose(res[0][6], np.array(chain), rtol=1e-6, atol=1e-8)
