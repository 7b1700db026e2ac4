#!/usr/bin/env python
"""Call latency of small traces through the host-buffer C-ABI call (what ray
aiming issues: System.aim_chief / aim_marginal trace 1-3 rays hundreds of times,
rayopt/system.py:507-555), zero-copy path vs the H2D / kernel / D2H path."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from conftest import load_golden
from rayopt_b200.engine import Engine

c = load_golden("cooke_single_ray")
for mode in ("zero-copy", "copy-engines"):
    if mode == "copy-engines":
        os.environ["RTX_NO_ZERO_COPY"] = "1"
    eng = Engine(0)
    for n in (1, 3, 13, 100, 1000):
        y = np.repeat(c["y0"], n, 0)
        u = np.repeat(c["u0"], n, 0)
        for _ in range(50):
            eng.trace(c["table"], y, u)
        t0 = time.perf_counter()
        for _ in range(1000):
            eng.trace(c["table"], y, u)
        dt = (time.perf_counter() - t0)/1000*1e6
        t0 = time.perf_counter()
        for _ in range(1000):
            eng.trace(c["table"], y, u, keep_last=True, want=("y",))
        dl = (time.perf_counter() - t0)/1000*1e6
        print("%-12s N=%5d rays, S=8: %6.1f us per full-trace call, %6.1f us keep-LAST y only" % (mode, n, dt, dl), flush=True)
    eng.close()
