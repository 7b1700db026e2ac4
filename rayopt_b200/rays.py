"""Launch-ray generation for aimed bundles (host numpy restatements; the device
kernels rtx_aim_infinite / rtx_aim_finite follow them operation by operation).

Restates, for the cases the benchmark workloads need, what the reference does
upstream of the hot path:  ``InfiniteConjugate.aim`` with the rectilinear
projection (rayopt/conjugates.py:208-213, 236-255), ``Pupil.map`` with
``filter=False`` (rayopt/pupils.py:97-107) and ``sagittal_meridional``
(rayopt/utils.py:102-114), given a pupil-aiming solution ``(z, p)`` that the
reference's ``System.pupil`` solved on the host (rayopt/system.py:585-593).
The object surface (``system[0]``) is taken to be a plane, as in every fixture.
"""
import numpy as np


def disc(n, seed):
    """uniform pupil coordinates in the unit disc: r = sqrt(U), phi = 2 pi U"""
    rng = np.random.default_rng(seed)
    r = np.sqrt(rng.random(n))
    phi = 2*np.pi*rng.random(n)
    return np.c_[r*np.cos(phi), r*np.sin(phi)]


def hexapolar(nrays):
    """pupil_distribution("hexapolar", nrays) (rayopt/utils.py:174-180):
    returns (rings, xy (N,2)); N = 1 + 3 rings (rings + 1)"""
    n = int(np.sqrt(nrays/3. - 1/12.) - 1/2.)
    l = [np.zeros((2, 1))]
    for i in range(1, n + 1):
        a = np.linspace(0, 2*np.pi, 6*i, endpoint=False)
        l.append([np.sin(a)*i/n, np.cos(a)*i/n])
    return n, np.concatenate(l, axis=1).T


def hexapolar_xy(idx, rings):
    """pupil coordinates of rays `idx` of the hexapolar grid with `rings`
    rings (rayopt/utils.py:174-180) without building the whole grid: ray 0 on
    axis; ring i = 1..rings holds rays 3 i (i-1) < j <= 3 i (i+1) at angles
    k 2 pi/(6 i) -- the index arithmetic of the device generator (pupil_xy in
    rtx_device.cuh), used to check samples of 1e8-ray device-generated bundles"""
    idx = np.asarray(idx, np.int64)
    i = np.floor((1 + np.sqrt(1 + 4*np.maximum(idx - 1, 0)/3.))/2).astype(np.int64)
    i = np.where(3*i*(i + 1) < idx, i + 1, i)
    i = np.where(3*i*(i - 1) >= idx, i - 1, i)
    i = np.maximum(i, 1)
    k = idx - 1 - 3*i*(i - 1)
    a = k*(2*np.pi/(6*i))
    xy = np.c_[np.sin(a)*i/rings, np.cos(a)*i/rings]
    xy[idx == 0] = 0
    return xy


def aim_frame(yo, z, angle):
    """The per-field constants of aim_infinite: (u, ybase, s, m), each (3,):
    the common ray direction, yz - z*u, and the normalised sagittal and
    meridional pupil axes."""
    yo = np.atleast_2d(np.asarray(yo, float))
    yt = yo*np.tan(angle)
    u = np.hstack((yt, np.ones((1, 1))))
    u /= np.sqrt(np.square(u).sum(-1))[:, None]
    yz = np.array((0, 0, z), float)
    ybase = yz - z*u
    s = np.cross(u, yz)
    if np.all(s == 0):
        s = np.array([[1., 0, 0]])
    m = np.cross(u, s)
    s = s/np.sqrt(np.square(s).sum(-1))[..., None]
    m = m/np.sqrt(np.square(m).sum(-1))[..., None]
    return u[0], ybase[0], s[0], m[0]


def aim_infinite(yo, yp, z, p, angle):
    """Rays (y, u), each (N, 3), for fractional object coordinate `yo` (2,),
    fractional pupil coordinates `yp` (N, 2), pupil distance `z`, pupil
    half-apertures `p` (2, 2) and object semi-angle `angle` (radians)."""
    yo = np.atleast_2d(np.asarray(yo, float))
    yp = np.atleast_2d(np.asarray(yp, float))
    p = np.asarray(p, float)
    yp = yp*np.fabs(p).max()                               # pupils.py:100-101
    yo, yp = np.broadcast_arrays(yo, yp)
    n = yo.shape[0]
    yt = yo*np.tan(angle)                                  # conjugates.py:211
    u = np.hstack((yt, np.ones((n, 1))))
    u /= np.sqrt(np.square(u).sum(-1))[:, None]
    yz = np.array((0, 0, z), float)
    y = yz - z*u                                           # conjugates.py:249
    s = np.cross(u, yz)                                    # utils.py:103-110
    axial = np.all(s == 0, axis=-1)[..., None]
    s = np.where(axial, (1., 0, 0), s)
    m = np.cross(u, s)
    s /= np.sqrt(np.square(s).sum(-1))[..., None]
    m /= np.sqrt(np.square(m).sum(-1))[..., None]
    y += yp[..., 0, None]*s + yp[..., 1, None]*m           # conjugates.py:252
    y += (-y[:, 2]/u[:, 2])[..., None]*u                   # plane object surface, :254
    return y, u


def finite_frame(yo, z, radius):
    """Per-field constants of aim_finite: object point y (3,), chief direction
    u = (0,0,z) - y (3,), normalised sagittal and meridional axes s, m (3,)."""
    yo = np.atleast_2d(np.asarray(yo, float))
    y = np.zeros((1, 3))
    y[..., :2] = -yo*radius                                # conjugates.py:151
    uz = np.array((0, 0, z), float)
    u = uz - y                                             # conjugates.py:158
    s = np.cross(u, uz)
    if np.all(s == 0):
        s = np.array([[1., 0, 0]])
    m = np.cross(u, s)
    s = s/np.sqrt(np.square(s).sum(-1))[..., None]
    m = m/np.sqrt(np.square(m).sum(-1))[..., None]
    return y[0], u[0], s[0], m[0]


def aim_finite(yo, yp, z, p, radius):
    """FiniteConjugate.aim (rayopt/conjugates.py:137-166) for a non-telecentric
    pupil, a plane object surface and ``filter=False``: object point at
    ``-yo*radius``, pupil coordinates mapped through the pupil half-angles
    ``arctan2(p, z)`` (Pupil.map, rayopt/pupils.py:97-101), direction
    ``normalize(u + z tan(yp_x) s + z tan(yp_y) m)``, flipped if z < 0."""
    yp = np.atleast_2d(np.asarray(yp, float))
    a = np.arctan2(np.asarray(p, float), z)                # conjugates.py:146
    yp = yp*np.fabs(a).max()                               # pupils.py:100-101
    yp = z*np.tan(yp)                                      # conjugates.py:149
    y0, u0, s, m = finite_frame(yo, z, radius)
    n = yp.shape[0]
    y = np.broadcast_to(y0, (n, 3)).copy()
    u = np.broadcast_to(u0, (n, 3)).copy()
    u += yp[..., 0, None]*s + yp[..., 1, None]*m           # conjugates.py:161
    u /= np.sqrt(np.square(u).sum(-1))[..., None]          # normalize, utils.py:97-98
    if z < 0:
        u *= -1
    return y, u
