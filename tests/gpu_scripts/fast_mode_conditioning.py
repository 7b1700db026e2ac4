#!/usr/bin/env python
"""How far is the default fast mode from the reference on ALL rays of the
random wild systems, measured against the reference's own conditioning (its
response to 1..16-ulp perturbations of the launch ray)?  Prints per seed:
rays, fraction within 1e-10, the worst ratio err_fast / max(1e-10, response),
NaN-mask flips not explained by a perturbation flip."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import np_oracle
from test_gpu_random_systems import random_table, random_rays, euler
from rayopt_b200.engine import Engine

np.seterr(all="ignore")
eng = Engine(0)


def per_ray_err(a_list, b_list):
    """(N,) max over arrays/surfaces of the SURVEY 8d error; (N,) any NaN-mask flip"""
    err, flip = 0, False
    for a, b in zip(a_list, b_list):
        a = a.reshape(a.shape[0], a.shape[1], -1)
        b = b.reshape(a.shape)
        scale = np.maximum(np.nanmax(np.where(np.isfinite(b), np.abs(b), 0), axis=(1, 2), keepdims=True), 1.0)
        d = np.nan_to_num(np.abs(a - b)/np.maximum(np.abs(b), scale))
        err = np.maximum(err, d.max(axis=(0, 2)))
        flip = flip | (np.isnan(a) != np.isnan(b)).any(axis=(0, 2))
    return err, flip


for kind, base, nseed in (("analytic", 1000, 24), ("general", 2000, 16)):
    for seed in range(nseed):
        rng = np.random.default_rng(base + seed)
        S = int(rng.integers(2, 24 if kind == "analytic" else 16))
        table = random_table(rng, S, rotated=kind == "general", newton=kind == "general")
        rot0 = None
        if kind == "general":
            rot0 = euler(*rng.normal(0, .02, 3)) if seed % 4 == 0 else None
            n = int(rng.choice([300, 5000]))
        else:
            n = int(rng.choice([257, 2000, 40003]))
        y0, u0 = random_rays(rng, n)
        clip = bool(seed % 2)
        want = np_oracle.trace(table, y0, u0, clip=clip, rot0=rot0)
        got = eng.trace(table, y0, u0, clip=clip, rot0=rot0)
        e_fast, f_fast = per_ray_err(got, want)
        resp, f_pert = np.zeros(n), np.zeros(n, bool)
        prng = np.random.default_rng(7)
        eps = 2.0**-52
        for mag in (1, 4, 16, 16):
            y1 = y0*(1 + mag*eps*prng.choice([-1, 1], y0.shape))
            u1 = u0.copy()
            u1[:, :2] *= (1 + mag*eps*prng.choice([-1, 1], (n, 2)))
            u1[:, 2] = np.sqrt(1 - np.square(u1[:, :2]).sum(1))
            pert = np_oracle.trace(table, y1, u1, clip=clip, rot0=rot0)
            e, f = per_ray_err(pert, want)
            resp = np.maximum(resp, e/mag)
            f_pert |= f
        hi = np_oracle.trace(table, y0, u0, clip=clip, rot0=rot0, dtype=np.longdouble)
        e_hi, _ = per_ray_err([h.astype(np.float64) for h in hi], want)   # the reference's own float64 noise
        bound = np.maximum(1e-10, 64*np.maximum(resp, e_hi))
        ok = ~f_fast | f_pert
        ratio = e_fast/bound
        print("%s seed %2d S=%2d n=%5d: within 1e-10 %.4f | worst err %.2e | worst err/bound %.3f | unexplained flips %d | "
              "rays over bound %d" % (kind, seed, S, n, (e_fast <= 1e-10).mean(), e_fast.max(), ratio[~f_fast].max(initial=0),
                                      int((f_fast & ~f_pert).sum()), int((ratio[~f_fast] > 1).sum())), flush=True)
eng.close()
