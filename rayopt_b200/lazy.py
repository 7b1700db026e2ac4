"""Device-resident trace results with lazy host materialisation.

The full trace of 1e7 rays x 12 surfaces is 10 GB; PCIe moves it in 0.2 s while
the kernel needs 1.7 ms.  The consumers of ``GeometricTrace`` mostly read single
rows -- spot diagrams ``y[-1]``, ``i[-1]`` (rayopt/analysis.py:269-280), fans
``y[-1]``, ``y[0]``, ``u[0]`` (:231-245), ``rms`` ``y[i]``
(rayopt/geometric_trace.py:171-183), ``refocus`` ``y[at]``, ``i[at]`` (:82-99)
-- so the resident drop-in keeps ``y,u,i,t`` in HBM and hands out ``LazyRows``
objects that copy a surface row to the host the first time it is indexed (and
the whole array only for ``np.asarray``).

``ResidentMixin`` is the resident flavour of ``PropagateMixin``
(geometric_trace.py): ``bind(rayopt.GeometricTrace, resident=True)`` puts it in
front of the reference class; ``ResidentTrace`` is the standalone class with the
on-device ray generation on top.
"""
import numpy as np

from .engine import default_engine
from .surface_table import pack_system

# rows above this size live in page-locked host buffers (full-rate PCIe)
PINNED_ROW_BYTES = 1 << 20


class LazyRows:
    """numpy-like view of a device array (rows, ld, k...) restricted to the
    first `n` columns.  Indexing with a leading integer (or a slice of rows)
    downloads just those rows, once; like ``GeometricTrace.y[j]`` in the
    reference the result is a VIEW of the trace's storage: the host buffer of
    a row is reused, so a later ``propagate`` + access refreshes it in place.
    Item assignment writes through to the device."""

    def __init__(self, darray, n, dtype=np.float64):
        self._d = darray
        self._n = int(n)
        self.dtype = np.dtype(dtype)
        self.shape = (darray.shape[0], self._n) + tuple(darray.shape[2:])
        self.ndim = len(self.shape)
        self.size = int(np.prod(self.shape))
        self._rows = {}        # valid host copies
        self._bufs = {}        # host buffers (kept across invalidation)
        self.fetched_bytes = 0

    def __len__(self):
        return self.shape[0]

    def invalidate(self, rows=None):
        if rows is None:
            self._rows.clear()
        else:
            for r in rows:
                self._rows.pop(r, None)

    def _buffer(self, r):
        b = self._bufs.get(r)
        if b is None:
            shape = (self._n,) + self.shape[2:]
            nbytes = int(np.prod(shape))*self._d.dtype.itemsize
            eng = getattr(self._d, "engine", None)
            if nbytes >= PINNED_ROW_BYTES and hasattr(eng, "pinned_empty"):
                b = eng.pinned_empty(shape, self._d.dtype)
            else:
                b = np.empty(shape, self._d.dtype)
            self._bufs[r] = b
        return b

    def set_row(self, r, value):
        """host-side write of one row (e.g. launch rays), mirrored to the
        device; `value` broadcasts to (n, k) or to the padded (ld, k) row"""
        r = range(self.shape[0])[r]
        drow = self._d.rows(r)
        v = np.asarray(value)
        full = drow.shape[1:]
        if v.shape == full:                              # whole padded row
            v = np.ascontiguousarray(v, self._d.dtype)
            drow.upload(v)
            host = v[:self._n]
        else:
            host = np.ascontiguousarray(np.broadcast_to(v, (self._n,) + self.shape[2:]),
                                        self._d.dtype)
            drow.upload(host)                            # the first n columns are contiguous
        b = self._buffer(r)
        b[...] = host
        self._rows[r] = b if b.dtype == self.dtype else b.astype(self.dtype)

    def row(self, r):
        r = range(self.shape[0])[r]
        a = self._rows.get(r)
        if a is None:
            b = self._buffer(r)
            self._d.rows(r).download(out=b)              # first n columns of the row
            a = b if b.dtype == self.dtype else b.astype(self.dtype)
            self._rows[r] = a
            self.fetched_bytes += b.nbytes
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

    def __setitem__(self, idx, value):
        if not isinstance(idx, tuple):
            idx = (idx,)
        head, rest = idx[0], idx[1:]
        rows = [range(self.shape[0])[int(head)]] if isinstance(head, (int, np.integer)) \
            else list(range(self.shape[0])[head])
        value = np.asarray(value)
        for k, r in enumerate(rows):
            v = value if isinstance(head, (int, np.integer)) or value.ndim < self.ndim else value[k]
            if rest:
                host = np.array(self.row(r))
                host[rest] = v
            else:
                host = v
            self.set_row(r, host)

    def __array__(self, dtype=None, copy=None):
        a = np.stack([self.row(r) for r in range(self.shape[0])])
        return a if dtype is None else a.astype(dtype)


class ResidentMixin:
    """``allocate / rays_given / propagate / rms / refocus`` of GeometricTrace
    (rayopt/geometric_trace.py:37-99,171-183) with the results resident in HBM.
    Same signatures and attribute names; ``y, u, i, t`` are ``LazyRows``.
    ``propagate`` moves no ray data over PCIe; ``rms`` and ``refocus`` reduce on
    the device (8 doubles come back)."""

    engine = None
    exact = False
    resident = True
    # i[j] == u[j-1] bit for bit in unrotated systems (system.py:461-463): u and
    # i are then two row-shifted views of ONE (S+2, ld, 3) device buffer and the
    # kernel stores 56 instead of 80 bytes per ray-surface
    alias_incidence = True
    _dev = None

    def _engine(self):
        if self.engine is None:
            self.engine = default_engine()
        return self.engine

    # ---- a1
    def allocate(self, nrays):
        eng = self._engine()
        self.free()
        self.length = L = len(self.system)
        self.nrays = nrays
        self._ld = ld = (nrays + 63)//64*64
        d = {"y": eng.empty((L, ld, 3)), "t": eng.empty((L, ld))}
        if self.alias_incidence:
            d["uext"] = eng.empty((L + 1, ld, 3))
            d["u"], d["i"] = d["uext"].rows(1, L + 1), d["uext"].rows(0, L)
            self._i_alias = True
        else:
            d["u"], d["i"] = eng.empty((L, ld, 3)), eng.empty((L, ld, 3))
            self._i_alias = False
        self._dev = d
        self._wrap()
        self.n = np.empty(L)
        self.w = None
        self.ref = None
        self.l = 1.
        self._w_dev = None

    def _wrap(self):
        d, n = self._dev, self.nrays
        self.y, self.u = LazyRows(d["y"], n), LazyRows(d["u"], n)
        self.i, self.t = LazyRows(d["i"], n), LazyRows(d["t"], n)

    def _materialize_i(self):
        """a rotated element breaks i[j] == u[j-1]: give `i` its own array"""
        eng, d = self._engine(), self._dev
        own = eng.empty((self.length, self._ld, 3))
        own.copy_from(d["i"])
        eng.sync()
        d["i"] = own
        self._i_alias = False
        self.i = LazyRows(own, self.nrays)

    def free(self):
        """give the HBM back now (otherwise: when the object is collected)"""
        if self._dev:
            for a in self._dev.values():
                a.free()
        self._dev = None
        w = getattr(self, "_w_dev", None)
        if w is not None:
            w[1].free()
        self._w_dev = None

    # ---- a2
    def rays_given(self, y, u, l=None, w=None, ref=0):
        """rayopt/geometric_trace.py:49-70: launch rays into row 0 -- straight
        from the caller's arrays into HBM (one H2D per array when they are
        (N,3) float64; page-locked arrays copy at full PCIe rate)."""
        pos, dirn = np.broadcast_arrays(*np.atleast_2d(y, u))
        count, width = pos.shape
        if self._dev is None or self.nrays != count:
            self.allocate(count)
        self.l = self.system.wavelengths[0] if l is None else l
        self._w_default = w is None
        self.w = np.ones(count)/count if w is None else w
        self._w_dev = None
        self.ref = ref
        if width != 3:
            y0 = np.zeros((count, 3))
            u0 = np.zeros((count, 3))
            y0[:, :width] = pos
            u0[:, :width] = dirn
            if width < 3:                                  # assumes forward rays, :65-67
                u0[:, 2] = np.sqrt(1 - np.square(u0[:, :2]).sum(-1))
            pos, dirn = y0, u0
        d = self._dev
        d["y"].rows(0).upload(pos)
        d["u"].rows(0).upload(dirn)
        d["i"].rows(0).copy_from(d["u"].rows(0), count*24)     # i[0] = u[0], :68
        self._zero_t0()                                        # t[0] = 0, :70
        for a in (self.y, self.u, self.i, self.t):
            a.invalidate()
        self.n[0] = self.system.refractive_index(self.l, 0)

    def _zero_t0(self):
        eng, row = self._engine(), self._dev["t"].rows(0)
        if hasattr(eng, "memset"):
            eng.memset(row, 0)
        else:
            row.upload(np.zeros(row.shape[1:]))

    def _cache_system(self):
        """Trace.propagate, rayopt/raytrace.py:32-36"""
        for name in ("path", "track", "origins", "mirrored"):
            try:
                setattr(self, name, getattr(self.system, name))
            except AttributeError:
                pass

    # ---- a3
    def propagate(self, start=1, stop=None, clip=False):
        """rayopt/geometric_trace.py:72-80 on the resident rows"""
        self._cache_system()
        init = start - 1
        table, n, rot0 = pack_system(self.system, self.l, start, stop, n0=self.n[init])
        rows = len(table)
        if rows == 0:
            return
        if self._i_alias and (rot0 is not None or (table["flags"] & 1).any()):
            self._materialize_i()
        d = self._dev
        sl = (start, start + rows)
        self._engine().trace_device(
            table, d["y"].rows(init), d["u"].rows(init),
            d["y"].rows(*sl), d["u"].rows(*sl),
            None if self._i_alias else d["i"].rows(*sl), d["t"].rows(*sl),
            N=self.nrays, ld=self._ld, clip=clip, rot0=rot0, exact=self.exact)
        self.n[start:start + rows] = n
        touched = range(start, start + rows)
        for a in (self.y, self.u, self.t):
            a.invalidate(touched)
        # an aliased i[j] is u[j-1]: rows start+1 .. start+rows changed
        self.i.invalidate(range(start + 1, start + rows + 1) if self._i_alias else touched)

    # ---- reductions that never bring the rays to the host
    def _weights(self):
        """device copy of self.w, or None for the default 1/N weights"""
        if self.w is None or getattr(self, "_w_default", False):
            return None
        if self._w_dev is None or self._w_dev[0] is not self.w:
            if self._w_dev is not None:
                self._w_dev[1].free()
            self._w_dev = (self.w, self._engine().to_device(np.asarray(self.w, float)))
        return self._w_dev[1]

    def rms(self, i=-1, ref=None):
        """GeometricTrace.rms (rayopt/geometric_trace.py:171-183) on the
        resident intercepts (moment passes on the device, 64 bytes back)."""
        eng = self._engine()
        i = range(self.length)[i]
        ref_point = None if ref is None else self.y[i][ref, :2]
        return eng.rms(self._dev["y"].rows(i), self._weights(), N=self.nrays,
                       ref_point=ref_point)

    def refocus(self, at=-1):
        """GeometricTrace.refocus (rayopt/geometric_trace.py:82-99): the
        least-squares focus shift from device moments of y[at], i[at]
        (rtx_focus_moments), then the re-trace."""
        eng = self._engine()
        at = range(self.length)[at]
        shift = eng.refocus_shift(self._dev["y"].rows(at), self._dev["i"].rows(at),
                                  self._weights(), N=self.nrays)
        self.system[at].distance += shift
        self.propagate()


class ResidentTrace(ResidentMixin):
    """Standalone resident drop-in (no rayopt import needed) with launch rays
    generated in HBM."""

    def __init__(self, system, engine=None, exact=False, alias_incidence=True):
        self.system = system
        self.engine = engine or default_engine()
        self.exact = exact
        self.alias_incidence = alias_incidence
        self._dev = None

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
        self._w_default, self._w_dev = True, None
        self.ref = ref
        frame = np.concatenate(aim_frame(yo, z, angle))
        pmax = float(np.fabs(np.asarray(p, float)).max())
        d = self._dev
        eng.aim_infinite_into(d["y"].rows(0), d["u"].rows(0), count, rings, frame, pmax, yp)
        d["i"].rows(0).copy_from(d["u"].rows(0), count*24)   # i[0] = u[0] (geometric_trace.py:68)
        self._zero_t0()
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
