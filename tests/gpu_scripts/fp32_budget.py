#!/usr/bin/env python
"""FP32 error budget: the FP32 CUDA kernels and a float32 numpy evaluation of
the reference's own formulas (oracle/np_oracle.py, dtype=float32) against the
FP64 reference, with the comparator of SURVEY 8(d) (per-surface scale for
lengths, 1 for direction cosines).  Prints one line per case and array."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import np_oracle
from conftest import golden_names, load_golden, load_systems
from rayopt_b200.engine import Engine
from rayopt_b200.rays import aim_infinite, disc

np.seterr(all="ignore")
eng = Engine(0)


def err(a, b):
    """(max rel err per SURVEY 8d over entries finite in both, #mask flips, worst surface)"""
    a = np.asarray(a, np.float64)
    flip = np.isnan(a) != np.isnan(b)
    fin = ~np.isnan(a) & ~np.isnan(b)
    absb = np.where(np.isnan(b), 0, np.abs(b))
    scale = np.maximum(absb.reshape(len(b), -1).max(1), 1.0).reshape((-1,) + (1,)*(b.ndim - 1))
    e = np.where(fin, np.abs(a - b)/np.maximum(np.abs(np.where(fin, b, 1)), scale), 0)
    per_surface = e.reshape(len(b), -1).max(1)
    return float(e.max()), int(flip.sum()), int(per_surface.argmax())


def report(name, table, y0, u0, clip, rot0, want):
    got = eng.trace(table, y0, u0, clip=clip, rot0=rot0, dtype=np.float32)
    orc = np_oracle.trace(table, y0, u0, clip=clip, rot0=rot0, dtype=np.float32)
    for w, g, o, b in zip("yuit", got, orc, want):
        eg, fg, sg = err(g, b)
        eo, fo, so = err(o, b)
        print("%-26s %s  cuda %.2e (flips %4d, surf %2d)   numpy-f32 %.2e (flips %4d, surf %2d)  n=%d"
              % (name, w, eg, fg, sg, eo, fo, so, y0.shape[0]), flush=True)


for name in golden_names():
    c = load_golden(name)
    report(name, c["table"], c["y0"], c["u0"], c["clip"], c["rot0"], (c["Y"], c["U"], c["I"], c["T"]))
systems = load_systems()
for sysname, n in (("double_gauss", 200000), ("cooke_asph", 200000), ("zoom", 200000), ("cooke", 200000)):
    ent = systems[sysname]
    aim = ent["aim"][0][3]
    y0, u0 = aim_infinite(aim["field"], disc(n, 9), aim["z"], aim["p"], ent["object_angle"])
    want = np_oracle.trace(ent["tables"][0], y0, u0, clip=True)
    report(sysname + "_200k", ent["tables"][0], y0, u0, True, None, want)
eng.close()
