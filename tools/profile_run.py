#!/usr/bin/env python
"""Small driver for ncu: 4096 chains (geometric config), a few short mems_run launches."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402

path = 1 if (len(sys.argv) < 2 or sys.argv[1] == "tc") else 0
C = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
inner = int(sys.argv[3]) if len(sys.argv) > 3 else 16
pb = helpers.load_problem("geometric")
eng = helpers.engine(pb, C, path)
x0 = np.random.default_rng(1).uniform(pb["low"], pb["high"], size=(C, 3))
eng.init_chains(x0, scale0=0.128, seed=1)
for _ in range(3):
    smp, _ = eng.run(inner, 0, inner)
torch.cuda.synchronize()
print("ok", float(smp.mean()))
