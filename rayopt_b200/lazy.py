"""Device-resident trace results with lazy host materialisation.

The full trace of 1e7 rays x 12 surfaces is 10 GB; PCIe moves it in 0.2 s while
the kernel needs 1.7 ms.  The consumers of ``GeometricTrace`` mostly read single
rows -- spot diagrams ``y[-1]``, ``i[-1]`` (rayopt/analysis.py:269-280), fans
``y[-1]``, ``y[0]``, ``u[0]`` (:231-245), ``rms`` ``y[i]``
(rayopt/geometric_trace.py:171-183) -- so ``ResidentTrace`` keeps ``y,u,i,t`` in
HBM and hands out ``LazyRows`` objects that copy a surface row to the host the
first time it is indexed (and the whole array only for ``np.asarray``).
"""
import numpy as np

from .engine import default_engine
from .surface_table import pack_system


class LazyRows:
    """numpy-like read-only view of a device array (rows, ld, k...) restricted
    to the first `n` columns; indexing with a leading integer (or a slice of
    rows) downloads just those rows, once."""

    def __init__(self, darray, n, dtype=np.float64):
        self._d = darray
        self._n = int(n)
        self.dtype = np.dtype(dtype)
        self.shape = (darray.shape[0], self._n) + tuple(darray.shape[2:])
        self.ndim = len(self.shape)
        self._rows = {}
        self.fetched_bytes = 0

    def __len__(self):
        return self.shape[0]

    def invalidate(self, rows=None):
        if rows is None:
            self._rows.clear()
        else:
            for r in rows:
                self._rows.pop(r, None)

    def set_row(self, r, value):
        """host-side write of one row (launch rays), mirrored to the device"""
        v = np.ascontiguousarray(np.broadcast_to(value, self.shape[1:]), self._d.dtype)
        self._d.rows(r).upload(v)
        self._rows[r] = v.astype(self.dtype, copy=False)

    def row(self, r):
        r = range(self.shape[0])[r]
        a = self._rows.get(r)
        if a is None:
            full = self._d.rows(r).download()[0]
            a = np.ascontiguousarray(full[:self._n]).astype(self.dtype, copy=False)
            self._rows[r] = a
            self.fetched_bytes += full.nbytes
        return a

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        head, rest = idx[0], idx[1:]
        if isinstance(head, (int, np.integer)):
            a = self.row(int(head))
            return a[rest] if rest else a
        if isinstance(head, slice):
            rows = range(self.shape[0])[head]
            a = np.stack([self.row(r) for r in rows]) if len(rows) else \
                np.empty((0,) + self.shape[1:], self.dtype)
            return a[(slice(None),) + rest] if rest else a
        return np.asarray(self)[idx]

    def __array__(self, dtype=None, copy=None):
        a = np.stack([self.row(r) for r in range(self.shape[0])])
        return a if dtype is None else a.astype(dtype)


class ResidentTrace:
    """``GeometricTrace`` with device-resident results (same constructor,
    ``allocate / rays_given / propagate`` signatures and attribute names;
    rayopt/geometric_trace.py:37-80).  ``y, u, i, t`` are ``LazyRows``; ``n``,
    ``w``, ``ref``, ``l``, ``nrays`` are as in the reference.  ``propagate``
    moves no ray data over PCIe: the launch rays were uploaded by
    ``rays_given`` and sub-range traces start from the resident row."""

    def __init__(self, system, engine=None, exact=False):
        self.system = system
        self.engine = engine or default_engine()
        self.exact = exact
        self._dev = None

    def allocate(self, nrays):
        eng = self.engine
        self.free()
        self.length = len(self.system)
        self.nrays = nrays
        self._ld = (nrays + 63)//64*64
        L, ld = self.length, self._ld
        self._dev = {k: eng.empty((L, ld, 3)) for k in "yui"}
        self._dev["t"] = eng.empty((L, ld))
        self.y = LazyRows(self._dev["y"], nrays)
        self.u = LazyRows(self._dev["u"], nrays)
        self.i = LazyRows(self._dev["i"], nrays)
        self.t = LazyRows(self._dev["t"], nrays)
        self.n = np.empty(L)
        self.w = None
        self.ref = None
        self.l = 1.

    def free(self):
        if self._dev:
            for a in self._dev.values():
                a.free()
        self._dev = None

    def rays_given(self, y, u, l=None, w=None, ref=0):
        pos, dirn = np.broadcast_arrays(*np.atleast_2d(y, u))
        count, width = pos.shape
        if self._dev is None or self.nrays != count:
            self.allocate(count)
        self.l = self.system.wavelengths[0] if l is None else l
        self.w = np.full(count, 1./count) if w is None else w
        self.ref = ref
        y0 = np.zeros((self._ld, 3))
        u0 = np.zeros((self._ld, 3))
        y0[:count, :width] = pos
        u0[:count, :width] = dirn
        if width == 2:
            u0[:count, 2] = np.sqrt(1 - (u0[:count, 0]**2 + u0[:count, 1]**2))
        for name, v in (("y", y0), ("u", u0), ("i", u0)):
            self._dev[name].rows(0).upload(v)
            getattr(self, name).invalidate()
            getattr(self, name)._rows[0] = v[:count].copy()
        self._dev["t"].rows(0).upload(np.zeros(self._ld))
        self.t.invalidate()
        self.n[0] = self.system.refractive_index(self.l, 0)

    def rays_infinite(self, yo, z, p, angle, l=None, nrays=None, yp=None, ref=0):
        """Launch rays of an aimed bundle for an infinite conjugate generated
        directly in HBM (Engine.aim_infinite_device / rtx_aim_infinite): the
        hexapolar grid with about `nrays` rays, or the DEVICE pupil coordinates
        `yp` (N,2); (z, p) is the reference's pupil-aiming solution
        (System.pupil).  Nothing crosses PCIe."""
        from .rays import aim_frame
        eng = self.engine
        if yp is None:
            rings = int(np.sqrt(nrays/3. - 1/12.) - 1/2.)
            count = 1 + 3*rings*(rings + 1)
        else:
            rings, count = 0, yp.shape[0]
        if self._dev is None or self.nrays != count:
            self.allocate(count)
        self.l = self.system.wavelengths[0] if l is None else l
        self.w = np.full(count, 1./count)
        self.ref = ref
        frame = np.concatenate(aim_frame(yo, z, angle))
        pmax = float(np.fabs(np.asarray(p, float)).max())
        for dst in ("u", "i"):           # i[0] = u[0] (geometric_trace.py:68)
            eng.aim_infinite_into(self._dev["y"].rows(0), self._dev[dst].rows(0), count, rings,
                                  frame, pmax, yp)
        self._dev["t"].rows(0).upload(np.zeros(self._ld))
        for a in (self.y, self.u, self.i, self.t):
            a.invalidate()
        self.n[0] = self.system.refractive_index(self.l, 0)

    def rays_point(self, yo, wavelength=None, nrays=11, distribution="hexapolar",
                   filter=None, stop=None, clip=False):
        """GeometricTrace.rays_point (rayopt/geometric_trace.py:204-209) for a
        rayopt ``System``: the pupil is aimed by the reference on the host
        (``system.pupil``); for an infinite rectilinear object with a plane
        first surface and the hexapolar distribution the rays are then
        generated in HBM, otherwise by ``system.aim`` on the host and
        uploaded.  Traces with `clip`."""
        s = self.system
        l = s.wavelengths[0] if wavelength is None else wavelength
        z, p = s.pupil(yo, l=wavelength, stop=stop)
        obj = s.object
        on_device = (distribution == "hexapolar" and not obj.finite and nrays > 1 and
                     getattr(obj, "projection", "rectilinear") == "rectilinear" and
                     not getattr(s[0], "curvature", 0.) and
                     getattr(s[0], "aspherics", None) is None and filter in (None, False) and clip)
        if on_device:
            self.rays_infinite(yo, z, p, obj.angle, l=l, nrays=nrays)
        else:
            from rayopt.utils import pupil_distribution      # the reference's own helper
            ref, yp, weight = pupil_distribution(distribution, nrays)
            y, u = s.aim(yo, yp, z, p, filter=(not clip) if filter is None else filter)
            self.rays_given(y, u, l, weight, ref)
        self.propagate(clip=clip)

    def propagate(self, start=1, stop=None, clip=False):
        init = start - 1
        table, n, rot0 = pack_system(self.system, self.l, start, stop, n0=self.n[init])
        rows = len(table)
        if rows == 0:
            return
        d = self._dev
        self.engine.trace_device(
            table, d["y"].rows(init), d["u"].rows(init),
            d["y"].rows(start, start + rows), d["u"].rows(start, start + rows),
            d["i"].rows(start, start + rows), d["t"].rows(start, start + rows),
            N=self.nrays, ld=self._ld, clip=clip, rot0=rot0, exact=self.exact)
        self.n[start:start + rows] = n
        for a in (self.y, self.u, self.i, self.t):
            a.invalidate(range(start, start + rows))

    # reductions that never bring the rays to the host
    def rms(self, i=-1, ref=None):
        """GeometricTrace.rms (rayopt/geometric_trace.py:171-183) on the
        resident intercepts (two moment passes, 64 bytes back)."""
        eng = self.engine
        i = range(self.length)[i]
        w = None if self.w is None else eng.to_device(np.asarray(self.w, float))
        ref_point = None if ref is None else self.y[i][ref, :2]
        r = eng.rms(self._dev["y"].rows(i), w, N=self.nrays, ref_point=ref_point)
        if w is not None:
            w.free()
        return r
