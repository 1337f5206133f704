"""world_size-2 gloo tests (CPU) of the N > 1 path's host logic: the row partition, the additivity of the
packed statistics under the all-reduce, shard-independent draws (global-row Philox keys), and identical host
chains on every rank.  The device step is stood in for by the oracle here (CPU box); the same statements are
checked on GPUs by tests/test_gpu_multi.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partition():
    from boom_b200.distributed import shard_range
    for n in (0, 1, 7, 1000, 10_000_000):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in edges]
            assert all(s == n // world for s in sizes[:-1])     # Imputer.hpp:348-375: remainder to the last worker
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import boom_b200
    from boom_b200.distributed import shard_range
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        n, p = 3001, 7
        X, y, nt, beta_true = O.synth_binomial(n, p, 3, seed=5, max_trials=3)
        mix = O.logit_mixture()
        row0, row1 = shard_range(n, world, rank)
        h = boom_b200.host()
        slab = boom_b200.MvnModel(np.zeros(p), np.eye(p))
        spike = boom_b200.VariableSelectionPrior(p, 0.4)
        rng = boom_b200.RNG(99)          # same sampler seed on every rank
        beta = np.zeros(p)
        bits = [True] + [False] * (p - 1)
        trace = []
        for it in range(5):
            # this rank's rows only, Philox keyed by the global row
            xtx, xty, ss, _ = O.logit_step(X[row0:row1], y[row0:row1], nt[row0:row1], beta, 10, mix, 77, it, row_offset=row0)
            packed = torch.from_numpy(np.concatenate([xtx.ravel(), xty, [float(ss), 0, 0, 0]]))
            dist.all_reduce(packed)      # the one exchange step of an iteration
            packed = packed.numpy()
            full = O.logit_step(X, y, nt, beta, 10, mix, 77, it)
            d = np.sqrt(np.diag(full[0]))
            assert np.max(np.abs(packed[:p * p].reshape(p, p) - full[0]) / np.outer(d, d)) < 1e-13
            assert np.allclose(packed[p * p:p * p + p], full[1], rtol=1e-12, atol=1e-12)
            assert packed[p * p + p] == n
            inc, b = h.spike_slab_sweep(rng, packed[:p * p].reshape(p, p), packed[p * p:p * p + p], slab, spike, bits, 1, False)
            bits = [bool(v) for v in inc > 0.5]
            beta = b
            trace.append(np.concatenate([inc, b]))
        trace = torch.from_numpy(np.array(trace))
        gathered = [torch.zeros_like(trace) for _ in range(world)]
        dist.all_gather(gathered, trace)
        for g in gathered[1:]:
            assert torch.equal(g, gathered[0]), "host chains diverged across ranks"
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, "FAIL: %r" % (e,)))
    finally:
        dist.destroy_process_group()


def test_two_rank_statistics_and_identical_chains():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [out.get(timeout=240) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def _student_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    from boom_b200.distributed import shard_range
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        n, p = 2503, 6
        X, y, bt = O.synth_student(n, p, 3, seed=12)
        row0, row1 = shard_range(n, world, rank)
        beta, sigma, nu = bt * 0.9, 1.4, 3.5
        # the Student-t step's statistics [X'WX | X'Wy | n, y'Wy, sum w, sum log w] and the two pieces of the log likelihood
        # (row part, n) are additive over the shards; the weights depend on the GLOBAL row only
        xtwx, xtwy, sc, w = O.student_step(X[row0:row1], y[row0:row1], beta, sigma, nu, 31, 2, row_offset=row0)
        full = O.student_step(X, y, beta, sigma, nu, 31, 2)
        assert np.array_equal(w, full[3][row0:row1])
        packed = torch.from_numpy(np.concatenate([xtwx.ravel(), xtwy, sc]))
        dist.all_reduce(packed)
        packed = packed.numpy()
        d = np.sqrt(np.diag(full[0]))
        assert np.max(np.abs(packed[:p * p].reshape(p, p) - full[0]) / np.outer(d, d)) < 1e-13
        assert np.allclose(packed[p * p:p * p + p], full[1], rtol=1e-12, atol=1e-12)
        assert np.allclose(packed[p * p + p:], full[2], rtol=1e-12)
        ll = torch.tensor([O.student_loglike(X[row0:row1], y[row0:row1], beta, sigma, nu)], dtype=torch.float64)
        dist.all_reduce(ll)
        assert abs(ll.item() - O.student_loglike(X, y, beta, sigma, nu)) < 1e-9 * n
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, "FAIL: %r" % (e,)))
    finally:
        dist.destroy_process_group()


def test_two_rank_student_t_statistics_and_likelihood():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_student_worker, args=(r, 2, port, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [out.get(timeout=240) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
