#!/usr/bin/env python
"""bench.py -- Gibbs iterations/sec of the auxiliary-mixture hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One step = one Gibbs iteration of the named workload: device step (latent draws + X'WX, X'Wz on this
rank's rows) -> NCCL all-reduce of the packed statistics (N > 1) -> device->host copy of the statistics ->
host small-state step (inclusion sweep + Cholesky draw of beta, as in the reference) -> beta host->device.
Nothing is skipped or cached between steps (W changes every iteration).

The JSON line's top level is the HEADLINE workload (default C3, the config BASELINE.json's metric is quoted on):
  value  iterations/s with the rows adopted from device tensors (resident in HBM before the timed region),
         timed with CUDA events on the context's stream around the K steps, max over ranks.
  e2e    the same iterations/s through the public sampler surface (model.set_method(sampler);
         model.sample_posterior()) on a model built from HOST arrays, host wall clock around K steps;
         every step copies beta host->device and the statistics device->host.  The rows themselves are
         uploaded once by the first draw (they are the model's data, not a per-step input); that one-time
         copy is reported beside it (upload_once_*), not hidden.
  roofline      the dominant kernel of the workload, from CUDA events around every launch of that kernel
                inside the timed region (boomgpu option "timing").
  cpu_baseline  the UNMODIFIED reference (oracle/_ref/boom_ref_driver, built from /root/reference by
                oracle/build_ref.sh) on this box's host cores, on a bounded row sample of the same workload
                (kind "reference"), with all host threads and -- SURVEY 8 d3 -- with one worker; only if that binary
                did not travel to the box: the oracle's C port, one thread (kind "port").  The same for --impl reference.
  secondary     the other BASELINE.json configs under the same clock (value, ms_per_step, roofline, clocks each), the active-set
                runs and the Student-t sibling (student_t):
                N = 1: C1 (1000 iterations), C2, C5, C4;  N > 1: C4 and C5 (the multi-GPU configs).
                plus c3_active_set / c4_active_set: the spike-and-slab configs with the active-set statistics option.
                HBM-bound kernels also carry roofline.burst (the same kernel after a 2 s pause: the sustained figure runs
                under the board's power cap, see profiles/README.md).
  e2e_adapter   (N = 1) the BOOM-typed adapter on BOOM's own model classes beside the standalone classes on the same rows
                (oracle/_ref/boom_adapter_demo bench, row samples of C3's and C4's shapes): ms per Gibbs iteration each.
  multi_gpu_parity / selftest   (N > 1) before timing: the all-reduced statistics of one step over the N shards against the
                same step on ONE context holding all rows (rank 0 regenerates them), and the agreement of a short sharded
                chain with the one-GPU chain -- tests/test_gpu_multi.py's checks, under the driver's own launch.

Strong scaling: the workload's n is fixed; N ranks hold n/N rows each (BASELINE.json: "n=10M,p=500 at 1/2/4/8 GPU").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name: (kind, sampler, n, p, nonzero) -- BASELINE.json configs[0..4]
WORKLOADS = {
    "c1": ("logit", "auxmix", 100_000, 20, 5),
    "c2": ("poisson", "auxmix", 1_000_000, 50, 5),
    "c3": ("logit", "spike", 10_000_000, 500, 20),
    "c4": ("logit", "spike", 2_000_000, 4000, 40),
    "c5": ("logit", "auxmix", 200_000_000, 16, 5),
}
DESCR = {
    "c1": "BinomialLogitAuxmixSampler n=100k p=20 (BASELINE.json configs[0])",
    "c2": "PoissonRegressionAuxMixSampler n=1M p=50 (configs[1])",
    "c3": "BinomialLogitSpikeSlabSampler n=10M p=500, 20 true nonzeros (configs[2], the config the metric is quoted on)",
    "c4": "BinomialLogitSpikeSlabSampler n=2M p=4000 (configs[3])",
    "c5": "BinomialLogitAuxmixSampler n=200M p=16 (configs[4])",
}
SEED = 20261017
CHUNK = 50_000                 # rows per generator chunk: the data do not depend on the sharding
FP64_DMMA_PEAK_TFLOPS = 37.0   # measured on this pool's B200 (profiles/r01_microbench_fp64.jsonl, DMMA issue-rate test)
HBM_FALLBACK_GBS = 6650.0
# DRAM bytes per observation measured by ncu (read + write), keyed by (kernel, p); see profiles/README.md
NCU_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")


def ncu_traffic():
    try:
        with open(NCU_TRAFFIC_FILE) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def beta_true(kind, p, nonzero):
    import numpy as np
    b = np.zeros(p)
    b[0] = -1.0 if kind == "logit" else 0.5
    for j in range(1, min(p - 1, nonzero) + 1):
        b[j] = 0.5 if j % 2 else -0.5
    return b


def make_shard(kind, n, p, nonzero, row0, row1, dev):
    """Rows [row0, row1) of the synthetic data set, generated on the device chunk by chunk."""
    import torch
    rows = row1 - row0
    X = torch.empty((rows, p), dtype=torch.float64, device=dev)
    y = torch.empty(rows, dtype=torch.float64 if kind == "logit" else torch.int64, device=dev)
    bt = torch.tensor(beta_true(kind, p, nonzero), dtype=torch.float64, device=dev)
    g = torch.Generator(device=dev)
    for c in range(row0 // CHUNK, (row1 + CHUNK - 1) // CHUNK):
        g.manual_seed(SEED + c)
        c0, c1 = c * CHUNK, min(n, (c + 1) * CHUNK)
        xc = torch.empty((c1 - c0, p), dtype=torch.float64, device=dev).normal_(generator=g)
        if kind == "poisson":
            xc.mul_(0.3)
        xc[:, 0] = 1.0
        eta = xc @ bt
        if kind == "logit":
            yc = (torch.rand(c1 - c0, dtype=torch.float64, device=dev, generator=g) < torch.sigmoid(eta)).double()
        else:
            yc = torch.poisson(torch.exp(eta), generator=g).long()
        a, b = max(c0, row0), min(c1, row1)
        X[a - row0:b - row0] = xc[a - c0:b - c0]
        y[a - row0:b - row0] = yc[a - c0:b - c0]
        del xc, eta, yc
    aux = torch.ones(rows, dtype=torch.float64, device=dev)  # trials / exposure
    return X, y, aux


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [f.strip() for f in line.split(",")]))

    def window(self, t0, t1):
        """Summary of the samples taken in [t0, t1] (the sampler keeps running: one process serves every workload)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if t1 - t0 < 0.12:
            time.sleep(0.12)
        ok = [(t, r) for (t, r) in list(self.rows) if len(r) >= 7]
        rows = [r for (t, r) in ok if t0 <= t <= t1]
        note = None
        if not rows and ok:   # timed region shorter than the sampling period: the sample nearest to it
            rows = [min(ok, key=lambda tr: min(abs(tr[0] - t0), abs(tr[0] - t1)))[1]]
            note = "timed region shorter than the sampling period; nearest sample"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [nm for k, nm in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        out = {"sm_mhz": sm[len(sm) // 2], "sm_mhz_min": sm[0], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
               "samples": len(rows), "reasons": reasons}
        if note:
            out["note"] = note
        return out

    def close(self):
        if self.proc is not None:
            self.proc.terminate()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    except (OSError, ValueError):
        return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------- the reference arm
def run_oracle_port(kind, sampler, n, p, nonzero, steps, warmup):
    """Fallback CPU baseline when the compiled reference is not on this box: the oracle's C restatement of the imputation
    pass (one thread) + the host small-state step, on a bounded row sample (kind "port")."""
    import numpy as np

    import boom_b200
    from oracle import oracle as O
    per_row_us = (0.6 if kind == "logit" else 1.5) + 0.00026 * p * p
    rows = int(min(n, max(2_000, 1.0e6 / per_row_us)))
    h = boom_b200.host()
    rng = boom_b200.RNG(SEED)
    slab = boom_b200.MvnModel(np.zeros(p), np.eye(p))
    spike = boom_b200.VariableSelectionPrior(p, min(1.0, max(nonzero, 1) / p))
    if kind == "logit":
        X, y, aux, _ = O.synth_binomial(rows, p, nonzero, SEED)
        mix = O.logit_mixture()
    else:
        X, y, aux, _ = O.synth_poisson(rows, p, nonzero, SEED)
        tab = O.poisson_table()
    beta = np.zeros(p)
    bits = [True] + [False] * (p - 1)
    t0 = None
    for it in range(warmup + steps):
        if it == warmup:
            t0 = time.perf_counter()
        if kind == "logit":
            xtx, xty, _, _ = O.logit_step(X, y, aux, beta, 10, mix, 1, it)
        else:
            xtx, xty, _ = O.poisson_step(X, y, aux, beta, tab, 1, it)
        if sampler == "spike":
            inc, beta = h.spike_slab_sweep(rng, xtx, xty, slab, spike, bits, 1, kind != "logit")
            bits = [bool(v) for v in inc > 0.5]
        else:
            beta = h.rmvn_suf(rng, xtx + np.eye(p), xty)
    secs = time.perf_counter() - t0
    ips = steps / secs
    return {"iters_per_sec": ips, "sample_rows": rows, "cores": 1, "kind": "port", "iters_per_sec_at_n": ips * rows / n}, None


def run_reference(kind, sampler, n, p, nonzero, steps, warmup, sample_rows=None, threads=None):
    """The unmodified reference on the host cores, on the first sample_rows rows' worth of the workload
    (the oracle port, one thread, when the compiled reference did not travel to this box)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "boom_ref_driver")
    if not os.path.exists(exe):
        return run_oracle_port(kind, sampler, n, p, nonzero, steps, warmup)
    cores = threads or os.cpu_count() or 1
    if sample_rows is None:
        # ~1-2 s per reference iteration on 16 cores: per-row cost ~ (0.5 + 0.26 p^2 / 1000) us single threaded
        per_row_us = 0.6 + 0.00026 * p * p
        sample_rows = int(min(n, 2_000_000, max(20_000, 1.5e6 * cores / per_row_us)))
    mode = {"auxmix": "logit", "spike": "spike"}[sampler] if kind == "logit" else "poisson"
    cmd = [exe, "bench", mode, str(sample_rows), str(p), str(nonzero), str(cores), str(steps), str(warmup)]
    if sampler == "spike":
        cmd.append("zellner")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=3000)
    if out.returncode != 0:
        return None, "reference driver failed: " + out.stderr.strip()[-200:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    r["sample_rows"] = sample_rows
    r["cores"] = cores
    r["kind"] = "reference"
    # per-row cost is constant (Imputer.hpp:177-179) and the host small-state step is negligible beside it,
    # so iterations/s at the workload's n = iterations/s on the sample * sample_rows / n
    r["iters_per_sec_at_n"] = r["iters_per_sec"] * sample_rows / n
    return r, None


GENERATOR_NOTE = ("rows from the oracle's bo_synth_* generator (same distribution as this arm's torch generator, different draws)")


def prior_text(sampler, nonzero, p):
    if sampler == "spike":
        return ("Zellner slab as LogitZellnerPrior builds it (priors.py:385-462): precision X'X/n with the off-diagonal halved, "
                "mean (trimmed logit of mean y, 0, ...); spike pi_j = %d/%d" % (nonzero, p))
    return "N(0, I)"


# ---------------------------------------------------------------------------------------------- our arm
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: boom_b200 has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.clocks = ClockSampler(self.local_rank) if self.rank == 0 and not os.environ.get("BENCH_NO_CLOCK_SAMPLER") else None
        self.peaks, self.peak_src = measured_peaks()
        self.traffic = ncu_traffic()

    # ---- plumbing
    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def zellner(self, X, y, n, p):
        """LogitZellnerPrior's slab from the rows (set-up, not timed): X'X by the library GEMM, summed over the ranks."""
        import numpy as np
        torch = self.torch
        G = torch.zeros((p, p), dtype=torch.float64, device=self.dev)
        step = max(1, min(X.shape[0], (1 << 28) // max(p, 1)))
        for a in range(0, X.shape[0], step):     # chunks bound cuBLAS's workspace; the result is one p x p matrix
            xa = X[a:a + step]
            G.addmm_(xa.t(), xa)
        ys = y.sum().reshape(1).double()
        if self.world > 1:
            self.dist.all_reduce(G)
            self.dist.all_reduce(ys)
        G = (G + G.t()).mul_(0.5 / n)
        d = torch.diagonal(G).clone()
        G.mul_(0.5)
        torch.diagonal(G).copy_(d)
        ph = min(0.999, max(0.001, float(ys.item()) / n))
        mean = np.zeros(p)
        mean[0] = np.log(ph / (1 - ph))
        return mean, G.cpu().numpy()

    def build(self, model, kind, sampler, n, p, nonzero, prior_mean, prior_prec, active=False):
        import numpy as np

        import boom_b200
        from boom_b200 import distributed as shard
        if prior_prec is None:
            prior = boom_b200.MvnModel(np.zeros(p), np.eye(p))
        else:
            prior = boom_b200.MvnModel(prior_mean, prior_prec, True)
        rng = boom_b200.RNG(SEED)
        if sampler == "spike":
            model.drop_all()
            model.add(0)   # chains start with only the intercept (GlmCoefs(p, all=false), as R's InitializeCoefficients does)
            spike = boom_b200.VariableSelectionPrior(p, min(1.0, max(nonzero, 1) / p))
            s = boom_b200.BinomialLogitSpikeSlabSampler(model, prior, spike, 10, rng)
            s.set_active_set_statistics(bool(active))
        elif kind == "logit":
            s = boom_b200.BinomialLogitAuxmixSampler(model, prior, 10, rng)
        else:
            s = boom_b200.PoissonRegressionAuxMixSampler(model, prior, 1, rng)
        model.set_method(s)
        shard.attach(model, n, self.stream, self.dev, self.rank, self.world, native=not self.args.torch_allreduce)
        model.set_device_option("timing", 1)
        for name, value in self.args.option:
            model.set_device_option(name, value)
        return s

    # ---- one workload: device-resident leg (+ optional host-array leg)
    def run(self, name, steps, warmup, e2e, n_override=0, active=False):
        import numpy as np

        import boom_b200
        from boom_b200 import distributed as shard
        torch, dist = self.torch, self.dist
        kind, sampler, n, p, nonzero = WORKLOADS[name]
        if n_override:
            n = n_override
        world, rank, dev = self.world, self.rank, self.dev
        row0, row1 = shard.shard_range(n, world, rank)
        X, y, aux = make_shard(kind, n, p, nonzero, row0, row1, dev)
        torch.cuda.synchronize()
        if kind == "poisson":
            boom_b200.load_poisson_mixture_table()
        pm, pp = self.zellner(X, y, n, p) if sampler == "spike" else (None, None)
        cfg = {"workload": DESCR[name] + (" [rows overridden to %d]" % n if n_override else ""), "n": n, "p": p,
               "true_nonzeros": nonzero, "sampler": sampler, "prior": prior_text(sampler, nonzero, p),
               "parallelism": "rows sharded over %d GPU(s), one NCCL all-reduce of p*p+p+4 doubles per iteration" % world}

        Model = boom_b200.BinomialLogitModel if kind == "logit" else boom_b200.PoissonRegressionModel
        model = Model(p)
        model.adopt_device_data(row1 - row0, X.data_ptr(), p, y.data_ptr(), aux.data_ptr())
        smp = self.build(model, kind, sampler, n, p, nonzero, pm, pp, active)
        if active:
            cfg["statistics"] = ("ACTIVE-SET option (set_active_set_statistics): X'WX for the included columns + diagonal + X'Wz per "
                                 "iteration, further columns fetched on accepted adds; same chain as the full matrix")
        for _ in range(warmup):
            model.sample_posterior()
        self.barrier()
        model.kernel_timings(True)
        launches0 = model.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(self.stream)
        for _ in range(steps):
            model.sample_posterior()
        e1.record(self.stream)
        self.barrier()
        t1 = time.perf_counter()
        dev_ms = self.max_over_ranks(e0.elapsed_time(e1))
        launches = model.kernel_launches() - launches0
        tm = model.kernel_timings(False)
        clk = self.clocks.window(t0, t1) if self.clocks else None
        beta_end = np.array(model.Beta)
        nvars = int(np.count_nonzero(np.array(model.inc)))
        if world > 1:   # every rank runs the same host chain on the same all-reduced statistics
            b = torch.tensor(beta_end, device=dev)
            lo, hi = b.clone(), b.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert torch.equal(lo, hi), "ranks diverged"
        value = steps / (dev_ms * 1e-3)

        # roofline of the dominant kernel; per-rank rows, this rank's launches
        my_rows = row1 - row0
        per = {k: (v[0] / max(v[1], 1), v[1]) for k, v in tm.items()}
        if p > 64 and active:
            # one read of X per iteration bounds this form: 8 n (p + 2) algorithmic bytes against the HBM roof
            k_ms = per["syrk_dmma"][0]
            nbytes = 8.0 * my_rows * (p + 2)
            roof = {"kernel": "panel_dmma_kernel", "bound": "hbm", "achieved": nbytes / (k_ms * 1e-3) * 1e-9, "peak": self.peaks["hbm_gbs"],
                    "unit": "GB/s", "peak_source": self.peak_src, "algorithmic_bytes_per_launch": nbytes,
                    "columns_fetched_in_timed_region": int(smp.active_set_columns_fetched)}
        elif p > 64:
            # the SYRK of one step: the main grid + (p not a multiple of 128) the launch for the ragged last block column
            k_ms = per["syrk_dmma"][0] * per["syrk_dmma"][1] / steps
            flops = float(my_rows) * p * (p + 1) + 2.0 * my_rows * p   # weighted SYRK (upper triangle) + X'Wz, FMA = 2
            roof = {"kernel": "syrk_dmma_kernel", "launches_per_step": per["syrk_dmma"][1] / steps,
                    "bound": "tensor", "achieved": flops / (k_ms * 1e-3) * 1e-12,
                    "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "peak_source": "FP64 DMMA issue-rate microbenchmark on this pool "
                    "(profiles/r01_microbench_fp64.jsonl; cuBLAS DGEMM 8192^3 = 35.4); MEASURED_PEAKS.json has no FP64 entry",
                    "algorithmic_flops_per_launch": flops}
            # the whole device step against the same peak: all algorithmic flops of the iteration / all kernel time
            step_ms = sum(v[0] * v[1] for v in per.values()) / steps
            roof["whole_step_frac"] = (flops + 2.0 * my_rows * p) / (step_ms * 1e-3) * 1e-12 / FP64_DMMA_PEAK_TFLOPS
        else:
            k_ms = per["fused_small"][0]
            nbytes = 8.0 * my_rows * (p + 2)
            roof = {"kernel": "fused_tma_kernel", "bound": "hbm", "achieved": nbytes / (k_ms * 1e-3) * 1e-9, "peak": self.peaks["hbm_gbs"],
                    "unit": "GB/s", "peak_source": self.peak_src, "algorithmic_bytes_per_launch": nbytes}
            if p >= 40:   # at the ridge (report both)
                flops = float(my_rows) * p * (p + 1) + 4.0 * my_rows * p
                roof["fp64_frac"] = flops / (k_ms * 1e-3) * 1e-12 / FP64_DMMA_PEAK_TFLOPS
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["kernel_ms"] = k_ms
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the same
        # kernel on the same workload (profiles/ncu_traffic.json), scaled to this rank's rows; null when no capture matches
        tr = self.traffic.get("%s:%d" % (roof["kernel"], p))
        roof["traffic"] = tr["bytes_per_row"] * my_rows if tr else None
        roof["traffic_source"] = tr["source"] if tr else None
        roof["kernel_ms_per_step"] = {k: round(v[0] * v[1] / steps, 4) for k, v in per.items() if v[1]}
        # HBM-bound kernels that also keep the FP64 pipes busy pull the board to its power cap after ~60 ms of back-to-back
        # launches (profiles/README.md, "Burst vs sustained"): the timed region above is the sustained figure; after a 2 s
        # pause the first launches run at the full clock -- reported beside it, never instead of it
        # (the decision is taken on the max over ranks: the extra iterations all-reduce, every rank must run them or none)
        if roof["bound"] == "hbm" and not self.args.no_burst and self.max_over_ranks(k_ms) >= 3.0:
            key = "syrk_dmma" if p > 64 else "fused_small"
            self.torch.cuda.synchronize()
            time.sleep(2.0)
            model.kernel_timings(True)
            for _ in range(4):
                model.sample_posterior()
            self.torch.cuda.synchronize()
            tb = model.kernel_timings(False)
            if tb[key][1]:
                b_ms = tb[key][0] / tb[key][1]
                roof["burst"] = {"kernel_ms": b_ms, "achieved": roof["achieved"] * k_ms / b_ms, "frac": roof["frac"] * k_ms / b_ms,
                                 "how": "4 iterations after a 2 s pause (board below its power cap, SM clock at maximum)"}

        # e2e through the sampler surface on a model built from HOST arrays
        e2e_out = None
        if e2e:
            del model, smp
            Xh = X.cpu().numpy()
            yh = y.cpu().numpy()
            ah = aux.cpu().numpy()
            del X, y, aux
            torch.cuda.empty_cache()
            model = Model(p)
            model.borrow_host_data(Xh, yh, ah)   # the rows stay in this process's host arrays; uploaded by the first draw
            nbytes_up = Xh.nbytes + yh.nbytes + ah.nbytes
            smp = self.build(model, kind, sampler, n, p, nonzero, pm, pp, active)
            self.barrier()
            tu = time.perf_counter()
            model.sample_posterior()     # packs and uploads the rows (once), then the first iteration
            self.barrier()
            first = time.perf_counter() - tu
            for _ in range(max(warmup - 1, 0)):
                model.sample_posterior()
            self.barrier()
            ta = time.perf_counter()
            for _ in range(steps):
                model.sample_posterior()
            self.barrier()
            wall = self.max_over_ranks(time.perf_counter() - ta)
            e2e_out = {"value": steps / wall, "unit": "iter/s", "h2d_bytes_per_step": 8 * p * world,
                       "d2h_bytes_per_step": 8 * (p * p + p + 4) * world, "timer": "host wall clock, max over ranks",
                       "upload_once_bytes": nbytes_up * 1, "first_iteration_with_upload_s": self.max_over_ranks(first)}
            del Xh, yh, ah
        del model, smp
        X = y = aux = None
        torch.cuda.empty_cache()

        in_bytes = 8.0 * my_rows * (p + 2)
        if in_bytes > 126e6:
            cfg["timing"] = "inputs (%.1f GB per GPU) larger than L2; CUDA events on the context stream; max over ranks" % (in_bytes * 1e-9)
        else:   # C1 only: the rows of a small model stay L2-resident from one Gibbs iteration to the next in real use as well
            cfg["timing"] = ("inputs (%.1f MB) SMALLER than the 126 MB L2 and not flushed: they are L2-resident between the iterations "
                             "of a real chain too; CUDA events on the context stream; max over ranks" % (in_bytes * 1e-6))
        out = {"value": value, "unit": "iter/s", "steps": steps, "warmup": warmup, "ms_per_step": dev_ms / steps, "config": cfg,
               "obs_per_sec": value * n, "gpu_launches": int(launches) * world, "clocks": clk, "roofline": roof,
               "model_size_at_end": nvars}
        if e2e_out is not None:
            out["e2e"] = e2e_out
        return out

    # ---- the Student-t sibling: TRegressionSampler on y = x'beta + 1.5 t_4, n = 25 M, p = 16 (C5's per-GPU shape)
    def run_student(self, steps, warmup, n=25_000_000, p=16):
        import numpy as np

        import boom_b200
        from boom_b200 import distributed as shard
        torch = self.torch
        world, rank, dev = self.world, self.rank, self.dev
        row0, row1 = shard.shard_range(n, world, rank)
        rows = row1 - row0
        g = torch.Generator(device=dev)
        g.manual_seed(SEED + 7919 * (rank + 1))
        X = torch.empty((rows, p), dtype=torch.float64, device=dev).normal_(generator=g)
        X[:, 0] = 1.0
        bt = torch.tensor(beta_true("poisson", p, 5), dtype=torch.float64, device=dev)
        z = torch.empty(rows, dtype=torch.float64, device=dev).normal_(generator=g)
        w = torch.empty(rows, dtype=torch.float64, device=dev).normal_(generator=g) ** 2
        for _ in range(3):
            w += torch.empty(rows, dtype=torch.float64, device=dev).normal_(generator=g) ** 2   # chi-square(4)
        y = X @ bt + 1.5 * z / torch.sqrt(w / 4.0)
        del z, w
        model = boom_b200.TRegressionModel(p)
        model.adopt_device_data(rows, X.data_ptr(), p, y.data_ptr())
        smp = boom_b200.TRegressionSampler(model, boom_b200.MvnModel(np.zeros(p), 100.0 * np.eye(p)), boom_b200.ChisqModel(1.0, 1.0),
                                           boom_b200.UniformModel(0.5, 60.0), boom_b200.RNG(SEED))
        model.set_method(smp)
        shard.attach(model, n, self.stream, dev, rank, world, native=True)
        model.set_device_option("timing", 1)
        for _ in range(warmup):
            model.sample_posterior()
        self.barrier()
        model.kernel_timings(True)
        launches0, evals0 = model.kernel_launches(), smp.likelihood_evaluations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(self.stream)
        for _ in range(steps):
            model.sample_posterior()
        e1.record(self.stream)
        self.barrier()
        t1 = time.perf_counter()
        dev_ms = self.max_over_ranks(e0.elapsed_time(e1))
        tm = model.kernel_timings(False)
        per = {k: (v[0] / max(v[1], 1), v[1]) for k, v in tm.items()}
        k_ms = per["fused_small"][0]
        nbytes = 8.0 * rows * (p + 2)
        roof = {"kernel": "fused_tma_kernel<kStudentT>", "bound": "hbm", "achieved": nbytes / (k_ms * 1e-3) * 1e-9, "peak": self.peaks["hbm_gbs"],
                "unit": "GB/s", "peak_source": self.peak_src, "algorithmic_bytes_per_launch": nbytes, "kernel_ms": k_ms,
                "kernel_ms_per_step": {k: round(v[0] * v[1] / steps, 4) for k, v in per.items() if v[1]}}
        roof["frac"] = roof["achieved"] / roof["peak"]
        out = {"value": steps / (dev_ms * 1e-3), "unit": "iter/s", "steps": steps, "warmup": warmup, "ms_per_step": dev_ms / steps,
               "config": {"workload": "TRegressionSampler (Student-t sibling, SURVEY 8 f4) n=%d p=%d, y = x'beta + 1.5 t_4" % (n, p), "n": n, "p": p,
                          "prior": "beta ~ N(0, 100 I), 1/sigma^2 ~ ChisqModel(1, 1), nu ~ U(0.5, 60)",
                          "iteration": "weights + WeightedRegSuf on the device (one pass over X), beta and sigma^2 on the host, nu by the slice "
                                       "sampler over the device log likelihood (one residual pass over X, then 8 n bytes per candidate nu)"},
               "obs_per_sec": steps / (dev_ms * 1e-3) * n, "gpu_launches": int(model.kernel_launches() - launches0) * world,
               "likelihood_evaluations_per_iteration": (smp.likelihood_evaluations - evals0) / steps,
               "clocks": self.clocks.window(t0, t1) if self.clocks else None, "roofline": roof,
               "posterior_draw_at_end": {"sigma": float(model.sigma), "nu": float(model.nu)}}
        del model, smp, X, y
        torch.cuda.empty_cache()
        return out

    # ---- N > 1: the sharded statistics against ONE context holding all rows, and a short chain against the one-GPU chain
    def selftest(self, name):
        import numpy as np

        import boom_b200
        from boom_b200 import distributed as shard
        torch, dist = self.torch, self.dist
        kind, sampler, n, p, nonzero = WORKLOADS[name]
        world, rank, dev = self.world, self.rank, self.dev
        if kind != "logit":
            return None
        # bounded: at most 2 M rows x p <= 500, or 250 k rows of a wide model -- the geometry (shards, offsets, split-K) is what counts
        n = min(n, 2_000_000 if p <= 500 else 250_000)
        row0, row1 = shard.shard_range(n, world, rank)
        X, y, aux = make_shard(kind, n, p, nonzero, row0, row1, dev)
        beta = beta_true(kind, p, nonzero) * 0.9
        mu, sigma, weights = boom_b200.default_logit_mixture()
        ctx = boom_b200.Context(self.local_rank)
        ctx.set_stream(self.stream.cuda_stream)
        ctx.set_logit_mixture(mu, sigma, weights)
        ctx.adopt_binomial(row1 - row0, p, X.data_ptr(), p, y.data_ptr(), aux.data_ptr())
        ctx.set_row_offset(row0)
        box = [boom_b200.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0], world, rank)
        xtx, xty, ss = ctx.logit_step(beta, 10, SEED, 3)       # all-reduces natively inside the C ABI
        ll, g, h = ctx.binomial_loglike_derivs(beta)             # all-reduced as well
        ctx.comm_destroy()
        ctx.close()
        res = {"workload_rows": n, "p": p, "ranks": world}
        # a short sharded chain through the sampler surface; compared below with rank 0's one-GPU chain
        def chain(model, attach):
            s = boom_b200.BinomialLogitSpikeSlabSampler(model, boom_b200.MvnModel(np.zeros(p), np.eye(p)),
                                                        boom_b200.VariableSelectionPrior(p, min(1.0, nonzero / p)), 10, boom_b200.RNG(17))
            model.set_method(s)
            model.drop_all()
            model.add(0)
            if attach:
                shard.attach(model, n, self.stream, dev, rank, world, native=True)
            else:
                model.set_device(self.local_rank)
                model.set_stream(self.stream.cuda_stream)
            out = []
            for _ in range(4):
                model.sample_posterior()
                out.append(np.array(model.Beta))
            return np.array(out)
        m = boom_b200.BinomialLogitModel(p)
        m.adopt_device_data(row1 - row0, X.data_ptr(), p, y.data_ptr(), aux.data_ptr())
        sharded_chain = chain(m, True)
        del m, X, y, aux
        torch.cuda.empty_cache()
        self.barrier()
        if rank == 0:
            Xf, yf, af = make_shard(kind, n, p, nonzero, 0, n, dev)
            one = boom_b200.Context(self.local_rank)
            one.set_stream(self.stream.cuda_stream)
            one.set_logit_mixture(mu, sigma, weights)
            one.adopt_binomial(n, p, Xf.data_ptr(), p, yf.data_ptr(), af.data_ptr())
            rxtx, rxty, rss = one.logit_step(beta, 10, SEED, 3)
            rll, rg, rh = one.binomial_loglike_derivs(beta)
            one.close()
            d = np.sqrt(np.abs(np.diag(rxtx)))
            dh = np.sqrt(np.abs(np.diag(rh)))
            res.update({
                "statistics_normwise_err": float(np.max(np.abs(xtx - rxtx) / np.outer(d, d))),
                "xty_rel_err": float(np.max(np.abs(xty - rxty)) / np.max(np.abs(rxty))),
                "sample_size_equal": bool(ss == rss == n),
                "loglike_rel_err": float(abs(ll - rll) / abs(rll)),
                "gradient_rel_err": float(np.max(np.abs(g - rg)) / np.max(np.abs(rg))),
                "hessian_normwise_err": float(np.max(np.abs(h - rh) / np.outer(dh, dh))),
            })
            m1 = boom_b200.BinomialLogitModel(p)
            m1.adopt_device_data(n, Xf.data_ptr(), p, yf.data_ptr(), af.data_ptr())
            one_chain = chain(m1, False)
            # same draws (Philox keyed by the global row), same host stream: the chains differ by summation order only
            res["chain_max_abs_diff_4_iterations"] = float(np.max(np.abs(one_chain - sharded_chain)))
            res["chain_same_model"] = bool(np.array_equal(one_chain != 0, sharded_chain != 0))
            del m1, Xf, yf, af
            tol = {"statistics_normwise_err": 1e-12, "xty_rel_err": 1e-11, "loglike_rel_err": 1e-12, "gradient_rel_err": 1e-10,
                   "hessian_normwise_err": 1e-12, "chain_max_abs_diff_4_iterations": 1e-6}
            res["tolerances"] = tol
            res["passed"] = bool(res["sample_size_equal"] and res["chain_same_model"] and all(res[k] <= v for k, v in tol.items()))
        torch.cuda.empty_cache()
        self.barrier()
        return res


def adapter_block():
    """The BOOM-typed drop-in (boom_b200/boom_adapter: samplers derived from BOOM::PosteriorSampler on BOOM's own model
    classes, n heap-allocated Data objects) beside the standalone classes on the same rows, at row samples of C3's and C4's
    shapes: what the BOOM surface costs per Gibbs iteration on top of the device step.  The demo binary links the reference
    library as the adapter's HOST FRAMEWORK, so it exists only where the reference was compiled."""
    exe = os.path.join(ROOT, "oracle", "_ref", "boom_adapter_demo")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/boom_adapter_demo not built (needs the reference sources)"}
    out = {"how": "boom_adapter_demo bench n p nonzero iters burn: BinomialLogitSpikeSlabSampler through BOOM's "
                  "model->sample_posterior() vs the standalone classes, same rows, wall clock per iteration"}
    for tag, argv in (("p500", ["400000", "500", "20", "10", "3"]), ("p4000", ["40000", "4000", "40", "4", "2"])):
        try:
            r = subprocess.run([exe, "bench"] + argv, capture_output=True, text=True, timeout=600)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            out[tag] = json.loads(line[-1]) if r.returncode == 0 and line else {"unavailable": (r.stderr or "no output")[-200:]}
            if "adapter_ms_per_iter" in out[tag]:
                out[tag]["adapter_over_standalone"] = out[tag]["adapter_ms_per_iter"] / out[tag]["standalone_ms_per_iter"]
        except (OSError, subprocess.TimeoutExpired, ValueError) as e:
            out[tag] = {"unavailable": repr(e)[-200:]}
    return out


def cpu_baseline_block(kind, sampler, n, p, nonzero):
    r, why = run_reference(kind, sampler, n, p, nonzero, 3, 1)
    if r is None:
        return {"unavailable": why}
    cpu = {"value": r["iters_per_sec_at_n"], "unit": "iter/s", "cores": r["cores"], "kind": r["kind"],
           "sample": "first %d of %d rows, 3 iterations after 1 warm-up, %d worker threads, scaled linearly in n; %s" % (
               r["sample_rows"], n, r["cores"], GENERATOR_NOTE), "measured_iters_per_sec_on_sample": r["iters_per_sec"]}
    if r["kind"] == "reference" and r["cores"] > 1:   # also with one worker, on a sample 1/8 the size
        rows1 = max(5_000, r["sample_rows"] // 8)
        r1, _ = run_reference(kind, sampler, n, p, nonzero, 2, 1, sample_rows=rows1, threads=1)
        if r1 is not None:
            cpu["one_worker"] = {"value": r1["iters_per_sec_at_n"], "unit": "iter/s", "cores": 1,
                                 "sample": "first %d rows, 2 iterations after 1 warm-up, scaled linearly in n" % rows1}
    return cpu


def parse_option(text):
    name, _, value = text.partition("=")
    return name, int(value)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-array e2e leg (development aid)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-adapter", action="store_true", help="skip the BOOM-adapter leg (e2e_adapter)")
    ap.add_argument("--no-burst", action="store_true", help="skip the burst re-measurement of HBM-bound kernels (2 s pause + 4 iterations)")
    ap.add_argument("--no-secondary", action="store_true", help="headline workload only (development aid, profiling)")
    ap.add_argument("--no-selftest", action="store_true", help="skip the N > 1 parity check before timing")
    ap.add_argument("--torch-allreduce", action="store_true",
                    help="all-reduce through a torch.distributed hook instead of the library's own NCCL communicator")
    ap.add_argument("--active-set", action="store_true", help="headline workload with the active-set statistics option (development aid)")
    ap.add_argument("--rows", type=int, default=0, help="override n (development aid; the line then names the override)")
    ap.add_argument("--option", type=parse_option, action="append", default=[], help="boomgpu option name=value (development aid)")
    args = ap.parse_args()
    kind, sampler, n, p, nonzero = WORKLOADS[args.workload]
    if args.rows:
        n = args.rows
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = args.steps, max(args.warmup, 0)
    metric = "gibbs_iterations_per_sec"

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        cfg = {"workload": DESCR[args.workload] + (" [rows overridden to %d]" % n if args.rows else ""), "n": n, "p": p,
               "true_nonzeros": nonzero, "sampler": sampler, "prior": prior_text(sampler, nonzero, p),
               "parallelism": "the reference's own worker pool on all host threads (set_number_of_workers)"}
        r, why = run_reference(kind, sampler, n, p, nonzero, steps, warmup)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": why}))
            return 0
        v = r["iters_per_sec_at_n"]
        sample = "first %d of %d rows, %d iterations after %d warm-up, %d worker threads, scaled linearly in n; %s" % (
            r["sample_rows"], n, steps, warmup, r["cores"], GENERATOR_NOTE)
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": v, "unit": "iter/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg, "obs_per_sec": v * n,
            "cpu_baseline": {"value": v, "unit": "iter/s", "cores": r["cores"], "kind": r["kind"], "sample": sample,
                             "measured_iters_per_sec_on_sample": r["iters_per_sec"]},
            "e2e": {"value": v, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ------------------------------------------------------------------ our arm
    b = Bench(args)
    st = None
    if world > 1 and not args.no_selftest:
        st = b.selftest(args.workload)
    head = b.run(args.workload, steps, warmup, e2e=not args.no_e2e, n_override=args.rows, active=args.active_set)
    secondary = {}
    if not args.no_secondary and not args.rows:
        # (name, steps, warm-up, e2e leg).  C1: 1000 iterations.  e2e only where the host copy of the rows is small.
        plan = [("c1", 1000, 20, True), ("c2", 200, 10, True), ("c5", 20, 3, False), ("c4", 4, 3, False)] if world == 1 else \
               [("c4", 6, 3, False), ("c5", 40, 5, False)]
        for nm, k, w, e in plan:
            if nm == args.workload:
                continue
            secondary[nm] = b.run(nm, k, w, e2e=e)
        # the optional active-set form of the spike-and-slab configs: same chain, fewer flops
        secondary["c3_active_set"] = b.run("c3", 20, 25, e2e=False, active=True)
        secondary["c4_active_set"] = b.run("c4", 10, 45, e2e=False, active=True)
        # the Student-t sibling
        secondary["student_t"] = b.run_student(20, 5)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_block(kind, sampler, n, p, nonzero)
    adapter = None
    if rank == 0 and world == 1 and not args.no_secondary and not args.rows and not args.no_adapter:
        b.torch.cuda.empty_cache()
        adapter = adapter_block()
    if b.clocks:
        b.clocks.close()
    if rank == 0:
        line = {"metric": metric, "value": head["value"], "unit": "iter/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": head["config"], "obs_per_sec": head["obs_per_sec"], "gpu_launches": head["gpu_launches"],
                "clocks": head["clocks"], "roofline": head["roofline"], "model_size_at_end": head["model_size_at_end"]}
        if "e2e" in head:
            line["e2e"] = head["e2e"]
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if adapter is not None:
            line["e2e_adapter"] = adapter
        if secondary:
            line["secondary"] = secondary
        if st is not None:
            line["multi_gpu_parity"] = st
            line["selftest"] = {"passed": st.get("passed"), "what": "sharded statistics / log likelihood derivatives vs one context holding "
                                "all rows; 4-iteration sharded chain vs the one-GPU chain (tests/test_gpu_multi.py's checks)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        b.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
