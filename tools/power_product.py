import os, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/tools')
import bench, helpers
pb = helpers.load_problem("geometric")
def sampled(fn, seconds=4.0):
    s = bench.ClockSampler(0); s.start(); time.sleep(0.3)
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < seconds:
        fn(); torch.cuda.synchronize(); n += 1
    dt = time.perf_counter() - t0
    clk = s.stop()
    pw = [float(r[3]) for r in s.rows if len(r) > 3]
    return clk, float(np.median(pw[3:])), n, dt
for Cn in (1 << 20, 148 * 2 * 128, 148 * 128):
    eng = helpers.engine(pb, Cn, 1)
    x0 = np.tile(pb["x_opts"], (Cn // 5 + 1, 1))[:Cn]
    eng.init_chains(x0, scale0=0.128, seed=1)
    S = 16 if Cn > 100000 else 256
    eng.run(S, 0, S, keep_samples=False); torch.cuda.synchronize()
    clk, pw, n, dt = sampled(lambda: eng.run(S, 0, S, keep_samples=False))
    print(f"{os.environ.get('MEMS_LIB','default')[-12:]:12s} chains {Cn:8d}: {pw:6.0f} W  SM {clk['sm_mhz']:6.0f} MHz  {Cn * S * n / dt / 1e6:7.2f} M chain-steps/s {clk['reasons']}", flush=True)
    eng.close()

# the same surrogate pipeline without the Metropolis bookkeeping: mems_forward on 4 Mi points per call (log-likelihoods only)
n = 1 << 22
eng = helpers.engine(pb, 1024, 1)
xd = torch.as_tensor(np.tile(pb["x_opts"], (n // 5 + 1, 1))[:n], device="cuda")
eng.forward(xd, want_y=False); torch.cuda.synchronize()
clk, pw, k, dt = sampled(lambda: eng.forward(xd, want_y=False))
print(f"forward only, {n} points per call: {pw:6.0f} W  SM {clk['sm_mhz']:6.0f} MHz  {n * k / dt / 1e6:7.2f} M traces/s {clk['reasons']}", flush=True)
