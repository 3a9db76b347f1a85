#!/usr/bin/env python
"""Benchmark of the batched surrogate-MH hot path (BASELINE.json metric: surrogate-scored
chain-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload sweep_1M|geometric_4096]

One bench "step" = one pass of the hot path over one batch of synthetic input: ``--inner``
Metropolis-Hastings updates of every chain (one ``mems_run`` launch per GPU).

Workloads
  sweep_1M        (default) BASELINE.json configs[4]: 1 048 576 chains in TOTAL, sharded over the ranks by global chain id
                  (strong scaling); geometric problem, T = 150, the shipped surrogate weights, synthetic observation.
  geometric_4096  BASELINE.json configs[1]: 4096 chains PER GPU (weak scaling); measured as well in every default run
                  and reported under "geometric_4096" in the same JSON line.

Reported
  value      resident: inputs already in HBM, CUDA events around the K launches, max over ranks
  e2e        the drop-in call a notebook makes -- ``MH(target, proposal, x0=[C,P]).sample_adapt(N, 0, Nt)`` ->
             ``BatchedSamples`` -> ``.samples`` -- with x0 in pinned host memory and the kept draws read back to pinned
             host memory inside the timed region
  end_of_run the collectives of the sweep, timed on NCCL: R-hat / posterior-moment all-reduces, the all-gather of the kept
             draws, the ESS reduction, the rank-normalised (arviz-default) R-hat and bulk ESS over all ranks -- and their
             share of a full 10 000-update run
  roofline   tensor-core: algorithmic FLOPs of the surrogate rows actually evaluated (device counter) / launch time

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "surrogate-scored chain-steps/sec"
UNIT = "chain-steps/s"

# The contract is ONE JSON line on stdout.  Libraries write banners to file descriptor 1 behind Python's back (NCCL
# prints its version there when /etc/nccl.conf or the environment asks for it), so fd 1 is pointed at stderr for the
# whole run and the JSON line goes to a private duplicate of the original stdout.
_REAL_STDOUT = None


def _protect_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def flop_per_row(P):
    return 2 * ((P + 1) * 64 + 5 * 64 * 64 + 64)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["bf16_tflops"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), float(d["hbm_gbs"]), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


def measured_traffic(args, chains_per_gpu):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), when it was taken on
    this workload; None otherwise."""
    for name in ("r02_bench_kernel_traffic.json", "r01_bench_kernel_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(p):
            continue
        d = json.load(open(p))
        if (d.get("workload"), d.get("chains"), d.get("mh_updates_per_launch"), d.get("thin")) == \
                (args.workload, chains_per_gpu, args.inner, args.thin) and args.path == "tc":
            return int(d["dram_bytes_read"]) + int(d["dram_bytes_write"]), name
    return None, None


def fixture(args):
    """The inverse problem of the workload, rebuilt from the committed fixture (mcmc_mems_b200.fixtures)."""
    from mcmc_mems_b200 import fixtures
    return fixtures.load_fixture("geometric", n_time=args.time_points if args.time_points != 150 else 0)


def oracle_logd_of(fx):
    """CPU oracle of the same posterior (numpy restatement of utils.py:30-40 + cuqi distributions)."""
    from oracle import make_forward_model, make_posterior_logd
    p = fx.processor
    fwd = make_forward_model(p.time, p.scaler.mean_, p.scaler.scale_, fx.layers)
    return make_posterior_logd(fwd, fx.y_obs, fx.sigma2, fx.low, fx.high)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU arms
def _cpu_chain(args):
    """One reference-shaped chain: per-update numpy forward on [T,P+1] exactly as
    src/InverseProblems/utils.py:30-40 + the CUQIpy MH loop (oracle port)."""
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    T, n_updates, seed = args
    from types import SimpleNamespace
    from oracle import mh_sample
    fx = fixture(SimpleNamespace(time_points=150))           # (the resampled trace needs the GPU forward; CPU legs use T = 150)
    logd = oracle_logd_of(fx)
    L = np.linalg.cholesky(fx.cov)
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    mh_sample(logd, fx.x_opts[seed % len(fx.x_opts)], n_updates + 1, 0, 0.128, rng=rng, chol=L)
    return time.perf_counter() - t0


def cpu_single_chain_rate(T, seconds=12.0):
    """Oracle port, one chain, one thread, bounded sample."""
    import threadpoolctl
    with threadpoolctl.threadpool_limits(1):
        n = 2000
        total_n, total_t = 0, 0.0
        while total_t < seconds:
            total_t += _cpu_chain((T, n, total_n))
            total_n += n
    return total_n / total_t, total_n


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; TF/CUQIpy cannot be installed here),
    one independent chain per host core, rank 0 only.  Same config keys as our arm; each step is a bounded sample of
    the workload (``--ref-updates`` MH updates of one chain per core)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_updates = args.ref_updates
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            pool.map(_cpu_chain, [(args.time_points, max(50, n_updates // 10), 1000 + c) for c in range(cores)])
        t0 = time.perf_counter()
        for k in range(args.steps):
            pool.map(_cpu_chain, [(args.time_points, n_updates, k * cores + c) for c in range(cores)])
        dt = time.perf_counter() - t0
    value = args.steps * cores * n_updates / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.workload == "sweep_1M" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{cores} chains (one per host core) x {n_updates} MH updates per step (numpy restatement of "
                                   "utils.py:30-40 + CUQIpy 0.8.0 MH; TF/CUQIpy not installable in this image)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def chains_of(args, world):
    """(total chains, chains per GPU) of the workload."""
    if args.workload == "sweep_1M":
        total = args.chains if args.chains > 0 else 1 << 20
        return total, (total + world - 1) // world
    per = args.chains if args.chains > 0 else 4096
    return per * world, per


def workload_config(args, world):
    total, per = chains_of(args, world)
    ref = {"sweep_1M": "BASELINE.json configs[4]: synthetic sweep, 1M chains sharded over 1/2/4/8 B200 with NCCL R-hat/ESS gather "
                       "(geometric 3-parameter problem; a bench step is a block of the 10k updates)",
           "geometric_4096": "BASELINE.json configs[1]: Xaccelerometer_geometric, 3-param posterior, 4096 parallel chains on 1 B200"}
    return {"workload": args.workload, "reference_config": ref.get(args.workload, args.workload),
            "n_params": 3, "time_points": args.time_points, "chains_total": total, "chains_per_gpu": per,
            "mh_updates_per_step": args.inner, "thin": args.thin, "sigma2": 0.005,
            "proposal_scale": 0.128, "adaptation": "frozen", "x0": "LSQ optima + one proposal step of jitter",
            "surrogate": "[4]->64x6 tanh->1 (reference weights)",
            "l2": "flushed between timed steps (256 MiB write, untimed); working set is on-chip"}


# ------------------------------------------------------------------------------------------ GPU arm
def _x0_all(fx, total):
    """Chains start like the notebook's (identification.ipynb cells 12-14): at the least-squares optima, here
    jittered by one proposal step so that the chains are distinct."""
    rng = np.random.default_rng(1)
    Lc = np.linalg.cholesky(fx.cov)
    x0 = fx.x_opts[np.arange(total) % len(fx.x_opts)] + 0.128 * rng.standard_normal((total, 3)) @ Lc.T
    return np.clip(x0, fx.low, fx.high)


def measure(args, fx, eng_of, world, rank, dev, dist, total, per, S, thin, steps, warmup, clocks=None):
    """Resident timing of one workload on this rank's shard: returns dict(value, ms, launches, rows, wall, engine, ...)."""
    import torch
    from mcmc_mems_b200 import distributed as D
    c0, c1 = D.shard_range(total, rank, world)
    C = c1 - c0
    n_keep = S // thin
    x0_host = torch.from_numpy(np.ascontiguousarray(_x0_all(fx, total)[c0:c1])).pin_memory()      # [C][P]
    eng = eng_of(C)
    samples = torch.empty(n_keep, 3, C, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng.init_chains(x0_host.to(dev), scale0=0.128, seed=0x5EED0000, chain_id0=c0)
    for _ in range(warmup):                         # also lets the starts reach the posterior before timing
        eng.run_into(S, 0, thin, samples)
    sync_all()
    rows0 = eng.rows_evaluated()
    launches0 = eng.launches
    if clocks is not None:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    sync_all()
    wall0 = time.perf_counter()
    for k in range(steps):
        flush.fill_(k & 0xFF)                       # L2 flush, outside the event pair
        ev[k][0].record()
        eng.run_into(S, 0, thin, samples)
        ev[k][1].record()
    sync_all()
    wall = time.perf_counter() - wall0
    ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    rows = eng.rows_evaluated() - rows0
    del flush
    return {"value": float(total) * S * steps / (ms * 1e-3), "ms": ms, "launches": eng.launches - launches0, "rows": rows,
            "wall": wall, "eng": eng, "x0_host": x0_host, "samples": samples, "C": C, "c0": c0, "sync_all": sync_all,
            "max_over_ranks": max_over_ranks}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mcmc_mems_b200 import MH, Gaussian, MHEngine, PATH_FP32, PATH_TC, cuqi_compat, fixtures, surrogate_spec
    from mcmc_mems_b200 import distributed as D
    import __graft_entry__ as g

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if rank == 0:
        g.build()
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")    # an explicit setting wins over /etc/nccl.conf
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{local}"))
        dist.barrier()
    g.build()
    dev = torch.device(f"cuda:{local}")

    fx = fixture(args)
    T = len(fx.processor.time)
    path = PATH_TC if args.path == "tc" else PATH_FP32
    spec = surrogate_spec(fx.processor, fx.nn_model)            # NN_Model + data processor -> what the kernels need

    def eng_of(C):
        return MHEngine(spec, fx.y_obs, fx.sigma2, fx.low, fx.high, fx.cov, n_chains=C, path=path, device=local)

    total, per = chains_of(args, world)
    S, thin = args.inner, args.thin
    clocks = ClockSampler(local) if rank == 0 else None
    m = measure(args, fx, eng_of, world, rank, dev, dist, total, per, S, thin, args.steps, args.warmup, clocks)
    eng, C, c0 = m["eng"], m["C"], m["c0"]
    sync_all, max_over_ranks = m["sync_all"], m["max_over_ranks"]
    value, ms_resident = m["value"], m["ms"]
    chain_steps = float(total) * S * args.steps

    # ---- end to end through the drop-in: the call INTEGRATION.md tells a notebook to make (identification.ipynb:300-308)
    #      x0 in pinned host memory -> MH(...).sample_adapt(N, 0, Nt) -> BatchedSamples -> .samples in pinned host memory
    post = fixtures.posterior_of(fx)
    prop = Gaussian(np.zeros(3), fx.cov)
    n_e2e = max(1, min(args.steps, args.e2e_steps))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    d2h_bytes = 0

    def dropin(seed):
        mh = MH(target=post, proposal=prop, x0=m["x0_host"], scale=0.128, seed=seed, path=path, device=local, chain_id0=c0)
        out = mh.sample(S + 1, 0, Nt=thin)           # cuqi counts x0 as the first sample: N = S + 1 -> S updates
        host = out.samples                            # device transpose + pinned device-to-host copy of the kept draws
        return host, out.acc_rate

    dropin(1)                                         # warm-up: builds and caches the engine of this setup
    e2e_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_e2e)]
    sync_all()
    for k in range(n_e2e):
        flush.fill_(k & 0xFF)
        e2e_ev[k][0].record()
        host, acc_rate = dropin(100 + k)
        e2e_ev[k][1].record()
        d2h_bytes = host.nbytes + acc_rate.nbytes
    sync_all()
    ms_e2e = max_over_ranks(sum(a.elapsed_time(b) for a, b in e2e_ev))
    e2e_value = float(total) * S * n_e2e / (ms_e2e * 1e-3)
    clk = clocks.stop() if rank == 0 else None
    cuqi_compat.close_engines()
    del flush

    # ---- end-of-run collectives of the sweep, on the clock (NCCL when world > 1): what a full run does once after its
    #      10 000 updates -- kept draws of this bench: `post_keep` thinned draws per chain
    eng.init_chains(m["x0_host"].to(dev), scale0=0.128, seed=0x5EED0000, chain_id0=c0)
    eng.run_into(args.post_burn, 0, args.post_burn, None)
    post_keep = args.post_keep
    post_smp = torch.empty(post_keep, 3, C, dtype=torch.float64, device=dev)
    eng.run_into(post_keep * args.post_thin, 0, args.post_thin, post_smp)
    acc_rate_chain = float(eng.accept_count().double().mean().item() / (args.post_burn + post_keep * args.post_thin))
    sync_all()

    def timed(fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        t0 = time.perf_counter()
        a.record()
        out = fn()
        b.record()
        sync_all()
        return out, max_over_ranks(max(a.elapsed_time(b), 1e3 * (time.perf_counter() - t0)))

    (rhat, (mean, std)), ms_allreduce = timed(lambda: (D.split_rhat(post_smp), D.posterior_moments(post_smp)))
    gathered, ms_gather = timed(lambda: D.gather_samples(post_smp[: args.gather_keep]))
    (ess, _), ms_ess = timed(lambda: D.ess(post_smp, max_lag=post_keep))
    gather_bytes = int(gathered.numel() * 8)
    gathered_shape = list(gathered.shape)
    del gathered
    # arviz-default (rank-normalised) R-hat and bulk ESS over every chain of every rank: all-gather of the pool, the library's
    # own sort / rank / quantile / Geyer kernels on each GPU, all-reduces of the moments and lag sums
    from mcmc_mems_b200 import diagnostics_device as DD
    rank_rhat, ms_rank_rhat = timed(lambda: DD.rhat_rank(post_smp))
    rank_ess, ms_rank_ess = timed(lambda: DD.ess_bulk(post_smp, max_lag=post_keep // 2))

    # ---- the 4096-chains-per-GPU configuration of BASELINE.json configs[1], in the same run
    second = None
    if args.workload == "sweep_1M" and args.second_steps > 0:
        from types import SimpleNamespace
        eng.close()
        a2 = SimpleNamespace(**{**vars(args), "workload": "geometric_4096", "chains": 4096, "inner": 256, "thin": 64})
        t2, p2 = chains_of(a2, world)
        m2 = measure(a2, fx, eng_of, world, rank, dev, dist, t2, p2, 256, 64, args.second_steps, 3)
        second = {"value": m2["value"], "unit": UNIT, "ms_per_step": m2["ms"] / args.second_steps, "scaling": "weak",
                  "chains_per_gpu": p2, "mh_updates_per_step": 256, "steps": args.second_steps,
                  "rows_evaluated_frac": m2["rows"] / (float(m2["C"]) * 256 * args.second_steps * T),
                  "tflops_per_gpu": m2["rows"] * flop_per_row(3) / (m2["ms"] * 1e-3) / 1e12}
        eng = m2["eng"]

    if rank == 0:
        peak_burst, peak_sust, hbm, which = load_peaks()
        flop_step = T * flop_per_row(3)
        t_launch = ms_resident * 1e-3 / args.steps
        rows_launch = m["rows"] / args.steps                        # surrogate rows this rank's launches really evaluated
        achieved = rows_launch * flop_per_row(3) / t_launch / 1e12   # per GPU, one launch
        nominal = float(C) * S * flop_step / t_launch / 1e12
        cpu_rate, cpu_n = cpu_single_chain_rate(T, args.cpu_seconds) if (world == 1 and args.cpu_seconds > 0) else (None, 0)
        traffic, traffic_src = measured_traffic(args, per)
        full_run_ms = 10000.0 / S * (ms_resident / args.steps)       # the sweep's 10k updates at the measured rate
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_resident / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.workload == "sweep_1M" else "weak", "vs_baseline": None,
            "dtype": "f32 (fp16 hi+lo split operands, f32 accumulate)" if path == PATH_TC else "f32",
            "data": "synthetic", "config": workload_config(args, world),
            "impl": "ours", "path": args.path,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(m["x0_host"].numel() * 8),
                    "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": ms_e2e / n_e2e, "steps": n_e2e,
                    "call": "MH(target, proposal, x0=[C,P] pinned host).sample(N, 0, Nt) -> BatchedSamples.samples (pinned host)"},
            "gpu_launches": int(m["launches"]),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_burst, "unit": "TFLOP/s",
                         "frac": achieved / peak_burst, "traffic": traffic,
                         "traffic_unit": f"bytes of DRAM traffic per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/{traffic_src})",
                         "peak_source": f"{which} burst bf16",
                         "frac_of_sustained": achieved / peak_sust,
                         "rows_evaluated_frac": rows_launch / (float(C) * S * T),
                         "rows_note": "proposals outside the prior box are not evaluated; FLOPs count evaluated surrogate rows only "
                                      "(device counter mems_rows_evaluated)",
                         "nominal_tflops_if_every_row_counted": nominal,
                         "issued_mma_frac": (achieved * (3.0 * 5 * 64 * 64 + 5 * 16 * 64 + 32 * 64) / (5 * 64 * 64 + 4 * 64 + 64)) / peak_burst
                         if path == PATH_TC else 0.0,
                         "mlp_rows_per_s": rows_launch / t_launch,
                         "sample_writeback_gbs": (S // thin) * 3 * C * 8 / t_launch / 1e9,
                         "flop_per_chain_step": flop_step},
            "end_of_run": {"all_reduce_ms": ms_allreduce, "all_gather_ms": ms_gather, "ess_ms": ms_ess,
                           "rank_rhat_ms": ms_rank_rhat, "rank_ess_ms": ms_rank_ess,
                           "all_gather_bytes": gather_bytes, "gathered_shape": gathered_shape,
                           "kept_draws_per_chain": post_keep, "backend": "nccl" if world > 1 else "none (1 rank)",
                           "full_run_ms_at_measured_rate": full_run_ms,
                           "share_of_full_run": (ms_allreduce + ms_gather + ms_ess + ms_rank_rhat + ms_rank_ess) / full_run_ms},
            "clocks": clk,
            "wall_s_timed_region": m["wall"],
            "check": {"split_rhat": [float(v) for v in rhat.cpu()], "post_mean": [float(v) for v in mean.cpu()],
                      "post_std": [float(v) for v in std.cpu()], "accept_rate": acc_rate_chain,
                      "ess_thinned_draws": [float(v) for v in ess], "kept_draws": int(post_keep * total),
                      "rank_normalised_split_rhat": [float(v) for v in rank_rhat],
                      "bulk_ess": [float(v) for v in rank_ess]},
        }
        if second is not None:
            line["geometric_4096"] = second
        if cpu_rate is not None:
            line["cpu_baseline"] = {"value": cpu_rate, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"1 chain x {cpu_n} MH updates, numpy restatement of utils.py:30-40 + "
                                              "CUQIpy 0.8.0 MH on one host core (TF/CUQIpy not installable here)"}
        _emit(line)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sweep_1M", choices=["sweep_1M", "geometric_4096"])
    ap.add_argument("--chains", type=int, default=0, help="sweep_1M: chains in total (default 1048576); geometric_4096: chains per GPU (default 4096)")
    ap.add_argument("--inner", type=int, default=0, help="MH updates per bench step (default: 32 for sweep_1M, 256 for geometric_4096)")
    ap.add_argument("--thin", type=int, default=0, help="keep every thin-th update (default: 32 / 64)")
    ap.add_argument("--time-points", type=int, default=150)
    ap.add_argument("--path", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-updates", type=int, default=4000, help="reference arm: MH updates per chain per step")
    ap.add_argument("--e2e-steps", type=int, default=5, help="drop-in calls timed for the end-to-end figure")
    ap.add_argument("--second-steps", type=int, default=5, help="launches of the geometric_4096 configuration timed next to the sweep (0 = skip)")
    ap.add_argument("--post-burn", type=int, default=200)
    ap.add_argument("--post-keep", type=int, default=16, help="thinned draws per chain kept for the end-of-run diagnostics")
    ap.add_argument("--post-thin", type=int, default=10)
    ap.add_argument("--gather-keep", type=int, default=16, help="kept draws per chain that go through the all-gather")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.inner <= 0:
        args.inner = 32 if args.workload == "sweep_1M" else 256
    if args.thin <= 0:
        args.thin = 32 if args.workload == "sweep_1M" else 64
    _protect_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
