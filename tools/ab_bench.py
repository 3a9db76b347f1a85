#!/usr/bin/env python
"""A/B of product-library builds on the bench workload: python tools/ab_bench.py name=DEF1,DEF2 name2=... [-- bench args].
Each variant is built with __graft_entry__.build_variant (here, before gpurun) and selected with MEMS_LIB on the box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

args = sys.argv[1:]
extra = []
if "--" in args:
    i = args.index("--")
    args, extra = args[:i], args[i + 1:]
build_only = "--build-only" in args
args = [a for a in args if a != "--build-only"]
for spec in args:
    name, _, defs = spec.partition("=")
    lib = g.LIB if name == "default" else g.build_variant(name, [d for d in defs.split(",") if d])
    if build_only:
        continue
    env = dict(os.environ, MEMS_LIB=lib)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--cpu-seconds", "0", *extra],
                         env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(f"{name:14s} {defs:40s} value {d['value'] / 1e6:7.2f} M chain-steps/s  e2e {d['e2e']['value'] / 1e6:7.2f} M  "
              f"ms/step {d['ms_per_step']:.3f}  clocks {d['clocks']['sm_mhz']}", flush=True)
    except Exception as e:
        print(name, "FAILED", e, out.stdout[-300:], out.stderr[-600:], flush=True)
