#!/usr/bin/env python
"""Benchmark of the batched surrogate-MH hot path (BASELINE.json metric: surrogate-scored
chain-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One bench "step" = one pass of the hot path over one batch of synthetic input: ``--inner``
Metropolis-Hastings updates of ``--chains`` chains per GPU (one ``mems_run`` launch).  Workload
(default) = BASELINE.json configs[1]: Xaccelerometer_geometric, 3 parameters, T = 150, 4096 chains
per B200, the shipped surrogate weights, synthetic observation.  Chains shard
across ranks with no data-path collective ("scaling": "weak"); NCCL is used once after the timed
region for the cross-chain R-hat / posterior moments.

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
sys.path.insert(0, os.path.join(ROOT, "tests"))

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


def measured_traffic(args):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), when it was taken on
    this workload; None otherwise."""
    p = os.path.join(ROOT, "profiles", "r01_bench_kernel_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    if (d.get("workload"), d.get("chains"), d.get("mh_updates_per_launch"), d.get("thin")) != \
            (args.workload, args.chains, args.inner, args.thin) or args.path != "tc":
        return None
    return int(d["dram_bytes_read"]) + int(d["dram_bytes_write"])


def problem(name, T_override=None):
    import helpers
    pb = helpers.load_problem(name)
    if T_override and T_override != len(pb["time"]):
        from oracle import batched_forward
        pb["time"] = np.linspace(0.0, 1.49e-3, T_override)
        f = batched_forward(pb["time"], pb["scaler_mean"], pb["scaler_scale"], pb["layers"], pb["x_true"][None])[0]
        pb["y_obs"] = f + np.sqrt(pb["sigma2"]) * np.random.default_rng(0).standard_normal(T_override)
    return pb


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
    name, T, n_updates, seed = args
    import helpers
    from oracle import mh_sample
    pb = problem(name, T)
    logd = helpers.oracle_logd(pb)
    L = np.linalg.cholesky(pb["cov"])
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    mh_sample(logd, pb["x_opts"][seed % len(pb["x_opts"])], n_updates + 1, 0, 0.128, rng=rng, chol=L)
    return time.perf_counter() - t0


def cpu_single_chain_rate(name, T, seconds=12.0):
    """Oracle port, one chain, one thread, bounded sample."""
    import threadpoolctl
    with threadpoolctl.threadpool_limits(1):
        n = 2000
        total_n, total_t = 0, 0.0
        while total_t < seconds:
            total_t += _cpu_chain((name, T, n, total_n))
            total_n += n
    return total_n / total_t, total_n


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; TF/CUQIpy cannot be installed here),
    one independent chain per host core, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_updates = args.ref_updates
    P = 3 if args.workload.startswith("geometric") else 4
    name = "geometric" if P == 3 else "qfactor"
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            pool.map(_cpu_chain, [(name, args.time_points, max(50, n_updates // 10), 1000 + c) for c in range(cores)])
        t0 = time.perf_counter()
        for k in range(args.steps):
            pool.map(_cpu_chain, [(name, args.time_points, n_updates, k * cores + c) for c in range(cores)])
        dt = time.perf_counter() - t0
    value = args.steps * cores * n_updates / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, P, cores_note=f"{cores} independent chains, one per host core"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{cores} chains x {n_updates} MH updates per step (numpy restatement of "
                                   "utils.py:30-40 + CUQIpy 0.8.0 MH; TF/CUQIpy not installable in this image)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def workload_config(args, P, cores_note=None):
    cfg = {"workload": args.workload, "reference_config": "BASELINE.json configs[1]: Xaccelerometer_geometric, "
           "3-param posterior, 4096 parallel chains on 1 B200" if args.workload == "geometric_4096" else args.workload,
           "n_params": P, "time_points": args.time_points, "chains_per_gpu": args.chains,
           "mh_updates_per_step": args.inner, "thin": args.thin, "sigma2": 0.005 if P == 3 else 0.5,
           "proposal_scale": 0.128, "adaptation": "frozen", "x0": "LSQ optima + one proposal step of jitter", "surrogate": f"[{P + 1}]->64x6 tanh->1 (reference weights)",
           "l2": "flushed between timed steps (256 MiB write, untimed); working set is on-chip"}
    if cores_note:
        cfg["cpu"] = cores_note
    return cfg


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import helpers
    from mcmc_mems_b200 import PATH_FP32, PATH_TC, distributed as D
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

    P = 3 if args.workload.startswith("geometric") else 4
    pb = problem("geometric" if P == 3 else "qfactor", args.time_points)
    T = len(pb["time"])
    C, S, thin = args.chains, args.inner, args.thin
    n_keep = S // thin
    path = PATH_TC if args.path == "tc" else PATH_FP32
    eng = helpers.engine(pb, C, path, device=local)
    # chains start like the notebook's (identification.ipynb cell 12-14): at the least-squares optima,
    # here jittered by one proposal step so that the C chains are distinct
    rng = np.random.default_rng(1)
    Lc = np.linalg.cholesky(pb["cov"])
    x0_all = pb["x_opts"][np.arange(C * world) % len(pb["x_opts"])] + 0.128 * rng.standard_normal((C * world, P)) @ Lc.T
    x0_all = np.clip(x0_all, pb["low"], pb["high"])
    c0, c1 = D.shard_range(C * world, rank, world)
    x0_host = torch.from_numpy(np.ascontiguousarray(x0_all[c0:c1].T)).pin_memory()      # [P][C]
    seed = 0x5EED0000
    samples = torch.empty(n_keep, P, C, dtype=torch.float64, device=dev)
    out_host = torch.empty(n_keep, P, C, dtype=torch.float64).pin_memory()
    acc_host = torch.empty(C, dtype=torch.int32).pin_memory()
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

    # ---- resident timing: inputs already in HBM, one mems_run launch per step -----------------
    eng.init_chains(x0_host.to(dev).t().contiguous(), scale0=0.128, seed=seed, chain_id0=c0)
    # let the uniform starts reach the posterior before timing: the mix of in-box / out-of-box
    # proposals does not change the arithmetic (every row is computed), only the accept rate
    for _ in range(args.warmup):
        eng.run_into(S, 0, thin, samples)
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                       # L2 flush, outside the event pair
        ev[k][0].record()
        eng.run_into(S, 0, thin, samples)
        ev[k][1].record()
    sync_all()
    wall = time.perf_counter() - wall0
    launches = eng.launches - launches0
    ms_resident = sum(a.elapsed_time(b) for a, b in ev)
    ms_resident = max_over_ranks(ms_resident)
    chain_steps = float(C) * world * S * args.steps
    value = chain_steps / (ms_resident * 1e-3)

    # ---- end to end through the public API: pinned host -> device -> kernels -> pinned host ----
    e2e_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        e2e_ev[k][0].record()
        xd = x0_host.to(dev, non_blocking=True)                     # H2D of this step's inputs
        eng.init_chains(xd.t(), scale0=0.128, seed=seed + k, chain_id0=c0)
        eng.run_into(S, 0, thin, samples)
        out_host.copy_(samples, non_blocking=True)                  # D2H of this step's result
        acc_host.copy_(eng.state()["accept_count"], non_blocking=True)
        e2e_ev[k][1].record()
    sync_all()
    ms_e2e = max_over_ranks(sum(a.elapsed_time(b) for a, b in e2e_ev))
    e2e_value = chain_steps / (ms_e2e * 1e-3)
    clk = clocks.stop() if rank == 0 else None

    # ---- end-of-run collectives (outside the timed region): R-hat sufficient stats + moments ----
    eng.init_chains(x0_host.to(dev).t().contiguous(), scale0=0.128, seed=seed, chain_id0=c0)
    eng.run(1000, 0, 1, keep_samples=False)
    post, _ = eng.run(800, 0, 8)
    rhat = D.split_rhat(post)
    mean, std = D.posterior_moments(post)
    ess, _ = D.ess(post, max_lag=post.shape[0])          # device autocovariance sums + one all-reduce
    acc_rate = float(acc_host.double().mean().item() / S)

    if rank == 0:
        peak_burst, peak_sust, hbm, which = load_peaks()
        flop_step = T * flop_per_row(P)
        t_launch = ms_resident * 1e-3 / args.steps
        achieved = float(C) * S * flop_step / t_launch / 1e12      # per GPU, one launch
        cpu_rate, cpu_n = cpu_single_chain_rate("geometric" if P == 3 else "qfactor", T, args.cpu_seconds) \
            if (world == 1 and args.cpu_seconds > 0) else (None, 0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_resident / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (fp16 hi+lo split operands, f32 accumulate)" if path == PATH_TC else "f32",
            "data": "synthetic", "config": workload_config(args, P),
            "impl": "ours", "path": args.path,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x0_host.numel() * 8),
                    "d2h_bytes_per_step": int(out_host.numel() * 8 + acc_host.numel() * 4),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_burst, "unit": "TFLOP/s",
                         "frac": achieved / peak_burst, "traffic": measured_traffic(args),
                         "traffic_unit": "bytes of DRAM traffic per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r01_bench_kernel_traffic.json)",
                         "peak_source": f"{which} burst bf16",
                         "frac_of_sustained": achieved / peak_sust,
                         "issued_mma_frac": (achieved * (3.0 * 5 * 64 * 64 + 5 * 16 * 64 + 32 * 64) / (5 * 64 * 64 + (P + 1) * 64 + 64)) / peak_burst
                         if path == PATH_TC else 0.0,
                         "mlp_rows_per_s": value / world * T,
                         "sample_writeback_gbs": n_keep * P * C * 8 / t_launch / 1e9,
                         "flop_per_chain_step": flop_step},
            "clocks": clk,
            "wall_s_timed_region": wall,
            "check": {"split_rhat": [float(v) for v in rhat.cpu()], "post_mean": [float(v) for v in mean.cpu()],
                      "post_std": [float(v) for v in std.cpu()], "accept_rate_last_step": acc_rate,
                      "ess_thinned_draws": [float(v) for v in ess], "kept_draws": int(post.shape[0] * post.shape[2] * world)},
        }
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
    ap.add_argument("--workload", default="geometric_4096")
    ap.add_argument("--chains", type=int, default=4096, help="chains per GPU")
    ap.add_argument("--inner", type=int, default=256, help="MH updates per bench step")
    ap.add_argument("--thin", type=int, default=64)
    ap.add_argument("--time-points", type=int, default=150)
    ap.add_argument("--path", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-updates", type=int, default=4000, help="reference arm: MH updates per chain per step")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    _protect_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
