"""numpy restatement of the device ray generator (rtx_aim_plan / rtx_aim_rays).

TEST INFRASTRUCTURE ONLY (like np_oracle.py).  Evaluates a `rtx_aim` record
(rayopt_b200/rays.py aim_record, include/rtx.h) the way the CUDA kernels
aim_candidate / aim_map / aim_rays_kernel do, operation by operation, so that

* on CPU it can be pinned against the reference itself --
  ``pupil_distribution`` (rayopt/utils.py:118-199), ``Pupil.map``
  (rayopt/pupils.py:97-107), ``InfiniteConjugate.aim`` / ``FiniteConjugate.aim``
  (rayopt/conjugates.py:137-166, 208-255) -- tests/test_aim_oracle.py;
* on the GPU the device output is compared with it bit for bit
  (tests/test_gpu_aim.py).
"""
import numpy as np

import np_oracle

GRID_GIVEN, GRID_HEXAPOLAR, GRID_SQUARE, GRID_TRIANGULAR, GRID_RANDOM, GRID_LINES = range(6)
_M64 = (1 << 64) - 1


def u01(seed, ctr):
    """the device's counter-based uniform generator (splitmix64 finaliser)"""
    x = (int(seed) + 0x9E3779B97F4A7C15*(int(ctr) + 1)) & _M64
    x = ((x ^ (x >> 30))*0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27))*0x94D049BB133111EB) & _M64
    x ^= x >> 31
    return (x >> 11)*(1.0/9007199254740992.0)


def _linspace_at(a, b, m, k):
    k = np.asarray(k, np.int64)
    if a == b or m < 2:
        return np.full(k.shape, float(a))
    step = (b - a)/float(m - 1)
    v = k*step + a
    return np.where(k == m - 1, b, v)


def candidates(rec, yp=None):
    """(px, py, keep) of every candidate index, as aim_candidate evaluates them"""
    g, n = int(rec["grid"]), int(rec["n"])
    if g == GRID_GIVEN:
        yp = np.asarray(yp, float).reshape(-1, 2)
        return yp[:, 0].copy(), yp[:, 1].copy(), np.ones(len(yp), bool)
    if g == GRID_HEXAPOLAR:
        from rayopt_b200.rays import hexapolar_xy
        xy = hexapolar_xy(np.arange(1 + 3*n*(n + 1)), n)
        return xy[:, 0], xy[:, 1], np.ones(len(xy), bool)
    if g in (GRID_SQUARE, GRID_TRIANGULAR):
        idx = np.arange(n*n)
        ix, iy = idx//n, idx % n
        step = 2.0/float(n - 1)
        x = ix*step + (-1.0)
        y = iy*step + (-1.0)
        if g == GRID_TRIANGULAR:
            x = np.where(iy & 1, x + 2.0/n, x)
        keep = (x*x + y*y) <= 1.0
        return np.r_[0., x], np.r_[0., y], np.r_[True, keep]
    if g == GRID_RANDOM:
        seed = int(rec["seed"])
        r = np.array([u01(seed, 2*j) for j in range(1, n + 1)])
        phi = np.array([u01(seed, 2*j + 1) for j in range(1, n + 1)])
        q = np.sqrt(r)
        return np.r_[0., np.cos(2*np.pi*phi)*q], np.r_[0., np.sin(2*np.pi*phi)*q], \
            np.ones(n + 1, bool)
    px, py = [], []
    for seg, m in zip(rec["seg"], rec["seg_m"]):
        k = np.arange(int(m))
        px.append(_linspace_at(seg[0], seg[2], int(m), k))
        py.append(_linspace_at(seg[1], seg[3], int(m), k))
    px, py = np.concatenate(px), np.concatenate(py)
    return px, py, np.ones(len(px), bool)


def generate(rec, yp=None):
    """-> (y (N,3), u (N,3), pupil (N,2)) of the kept rays, in order"""
    rec = rec[0] if getattr(rec, "ndim", 0) else rec
    px, py, keep = candidates(rec, yp)
    pmax = float(rec["pmax"])
    qx, qy = px*pmax, py*pmax                                  # Pupil.map
    if int(rec["filter"]):
        c, d2 = rec["fc"], rec["fd2"]
        with np.errstate(all="ignore"):
            keep = keep & ((qx - c[0])*(qx - c[0])/d2[0] + (qy - c[1])*(qy - c[1])/d2[1] <= 1.0)
    px, py, qx, qy = px[keep], py[keep], qx[keep], qy[keep]
    f = np.asarray(rec["frame"], float).reshape(4, 3)
    n = len(px)
    if int(rec["conjugate"]) == 0:
        u = np.tile(f[0], (n, 1))
        y = f[1] + (qx[:, None]*f[2] + qy[:, None]*f[3])
        if int(rec["curved"]):
            s = np_oracle.intercept(rec["surface"], y, u)
            y = y + s[:, None]*u
        else:
            y = y + (-y[:, 2]/u[:, 2])[:, None]*u
    else:
        z = float(rec["z"])
        y = np.tile(f[0], (n, 1))
        tx, ty = z*np.tan(qx), z*np.tan(qy)
        u = f[1] + (tx[:, None]*f[2] + ty[:, None]*f[3])
        u = u/np.sqrt((u[:, 0]*u[:, 0] + u[:, 1]*u[:, 1]) + u[:, 2]*u[:, 2])[:, None]
        if z < 0:
            u = -u
    return y, u, np.c_[px, py]
