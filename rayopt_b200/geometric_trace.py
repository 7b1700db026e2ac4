"""GPU drop-in for rayopt's real-ray trace driver.

Mirrors ``rayopt.geometric_trace.GeometricTrace`` (rayopt/geometric_trace.py)
for the hot path: same constructor, same ``allocate / rays_given / propagate``
signatures, same result attributes ``y, u, i (S+1, N, 3)``, ``t (S+1, N)``,
``n (S+1,)``, ``w, ref, l, nrays`` as numpy arrays -- but ``propagate`` is ONE
CUDA launch through the C ABI (include/rtx.h) instead of the Python loop over
``System.propagate`` (rayopt/system.py:459-464).  There is no CPU fallback.

Two ways to use it:

* standalone ``GeometricTrace(system)``: `system` is a rayopt ``System`` or any
  sequence of elements exposing what ``surface_table.pack_system`` reads;
* ``bind(rayopt.GeometricTrace)`` returns a subclass of the reference class
  whose ``allocate``/``propagate`` are replaced, so that ``rays_point``,
  ``rays_clipping``, ``refocus``, ``opd``, ``Analysis`` ... keep working
  unchanged.  ``bind(..., resident=True)`` keeps the results
  in HBM instead (``y,u,i,t`` are ``LazyRows``, lazy.py): the consumers above
  read single rows, so only those cross PCIe.
"""
import numpy as np

from .engine import default_engine
from .lazy import ResidentMixin
from .surface_table import pack_system

# result arrays above this size are allocated page-locked so that the D2H of a
# full trace runs at PCIe rate
PINNED_THRESHOLD = 8 << 20


class PropagateMixin:
    """`allocate` and `propagate` of GeometricTrace on the GPU engine."""

    engine = None        # rayopt_b200.engine.Engine; default: process-wide
    dtype = np.float64   # float32 selects the FP32 kernels (results cast up)
    exact = False        # RTX_EXACT: bit-identical-to-numpy FP64 arithmetic
    clip_default = False

    def _engine(self):
        if self.engine is None:
            self.engine = default_engine()
        return self.engine

    def _empty(self, shape):
        # page-locked memory is owned by the arrays themselves (freed when the
        # last view dies, Engine.pinned_empty): a row kept by the caller stays
        # valid after the trace object is gone or re-allocated
        if int(np.prod(shape))*8 >= PINNED_THRESHOLD:
            return self._engine().pinned_empty(shape, np.float64)
        return np.empty(shape)

    # For unrotated systems the incidence array is redundant: i[j] (direction
    # arriving at surface j, system.py:461-463) is bit-for-bit u[j-1] (direction
    # leaving surface j-1).  `u` and `i` are then two views of ONE (S+2, N, 3)
    # buffer shifted by a row, the kernel does not store `i`, and a full trace
    # moves 56 instead of 80 bytes per ray-surface over HBM and PCIe.  The
    # first rotated element (or rotated start frame) switches to a real array.
    alias_incidence = True

    def allocate(self, nrays):
        """rayopt/geometric_trace.py:37-47"""
        self.length = len(self.system)          # Trace.allocate, raytrace.py:29
        self.nrays = nrays
        self.n = np.empty(self.length)
        self.y = self._empty((self.length, nrays, 3))
        if self.alias_incidence:
            self._uext = self._empty((self.length + 1, nrays, 3))
            self.u = self._uext[1:]
            self.i = self._uext[:-1]
            self._i_alias = True
        else:
            self.u = self._empty((self.length, nrays, 3))
            self.i = self._empty((self.length, nrays, 3))
            self._i_alias = False
        self.w = None
        self.ref = None
        self.l = 1.
        self.t = self._empty((self.length, nrays))

    def _materialize_i(self):
        i = self._empty((self.length, self.nrays, 3))
        i[:] = self.i
        self.i = i
        self._i_alias = False

    def _cache_system(self):
        """Trace.propagate, rayopt/raytrace.py:32-36"""
        for name in ("path", "track", "origins", "mirrored"):
            try:
                setattr(self, name, getattr(self.system, name))
            except AttributeError:
                pass

    def propagate(self, start=1, stop=None, clip=False):
        """rayopt/geometric_trace.py:72-80 -- one launch instead of S numpy
        passes.  Rows start..stop-1 of y,u,i,t,n are overwritten."""
        self._cache_system()
        init = start - 1
        table, n, rot0 = pack_system(self.system, self.l, start, stop,
                                     n0=self.n[init])
        rows = len(table)
        if rows == 0:
            return
        sl = slice(start, start + rows)
        eng = self._engine()
        alias = getattr(self, "_i_alias", False)
        if alias and (rot0 is not None or (table["flags"] & 1).any()):
            self._materialize_i()
            alias = False
        if np.dtype(self.dtype) == np.float64:
            out = {"y": self.y[sl], "u": self.u[sl], "t": self.t[sl]}
            if not alias:
                out["i"] = self.i[sl]
            eng.trace(table, self.y[init], self.u[init], clip=clip, rot0=rot0,
                      exact=self.exact, out=out, want=tuple(out))
        else:
            want = ("y", "u", "t") if alias else ("y", "u", "i", "t")
            Y, U, I, T = eng.trace(table, self.y[init], self.u[init], clip=clip,
                                   rot0=rot0, dtype=self.dtype, want=want)
            self.y[sl], self.u[sl], self.t[sl] = Y, U, T
            if not alias:
                self.i[sl] = I
        self.n[sl] = n


class GeometricTrace(PropagateMixin):
    """Standalone drop-in (no rayopt import needed)."""

    def __init__(self, system, engine=None, dtype=np.float64, exact=False,
                 alias_incidence=True):
        self.system = system
        self.engine = engine
        self.dtype = dtype
        self.exact = exact
        self.alias_incidence = alias_incidence

    def rays_given(self, y, u, l=None, w=None, ref=0):
        """Load launch rays into row 0 (semantics of rayopt/geometric_trace.py:
        49-70): `y`, `u` broadcast against each other, (N, 2) input is padded
        with z = 0 and a forward u_z = +sqrt(1 - u_x^2 - u_y^2); the incidence
        of row 0 is the launch direction; default weights 1/N."""
        pos, dirn = np.broadcast_arrays(*np.atleast_2d(y, u))
        count, width = pos.shape
        if getattr(self, "y", None) is None or self.y.shape[1] != count:
            self.allocate(count)
        self.l = self.system.wavelengths[0] if l is None else l
        self.w = np.full(count, 1./count) if w is None else w
        self.ref = ref
        y0, u0 = self.y[0], self.u[0]
        y0[:, :width] = pos
        y0[:, width:] = 0
        u0[:, :width] = dirn
        if width == 2:
            u0[:, 2] = np.sqrt(1 - (u0[:, 0]**2 + u0[:, 1]**2))
        self.i[0] = u0
        self.t[0] = 0
        self.n[0] = self.system.refractive_index(self.l, 0)

    def rms(self, i=-1, ref=None):
        """Weighted rms spot radius at surface `i` about the mean intercept or
        about ray `ref` (rayopt/geometric_trace.py:171-183).  Like the
        reference it does NOT mask NaN rays."""
        spot = self.y[i, :, :2]
        centre = spot.mean(0) if ref is None else spot[ref]
        d2 = ((spot - centre)**2).sum(1)
        weights = np.full(d2.shape, 1./d2.size) if self.w is None else self.w
        return float(np.sqrt(np.dot(d2, weights)))

    def refocus(self, at=-1):
        """Least-squares focus shift of element `at` and re-trace
        (rayopt/geometric_trace.py:82-99): minimise sum w |y + t u|^2 about the
        means, u = tan of the incidence angles, over the rays that arrive."""
        inc = self.i[at]
        slope = inc[:, :2]/inc[:, 2:]                # tanarcsin, utils.py:42-48
        ok = np.isfinite(slope).all(1)
        yy = self.y[at, ok, :2]
        uu = slope[ok]
        ww = np.ones(len(yy)) if self.w is None else self.w[ok]
        yy = yy - yy.mean(0)
        uu = uu - uu.mean(0)
        shift = -(ww[:, None]*yy*uu).sum()/(ww[:, None]*uu*uu).sum()
        self.system[at].distance += shift
        self.propagate()


    # ---- ray launch helpers of the reference (rayopt/geometric_trace.py:185-229),
    # for a `system` that offers rayopt's aiming interface (System.pupil / aim /
    # aim_chief, rayopt/system.py:504-593); ray aiming stays host Python
    def rays(self, yo, yp, wavelength, stop=None, filter=None, clip=False, weight=None, ref=0):
        """aim pupil coordinates `yp` from field point `yo`, trace
        (geometric_trace.py:195-202)"""
        z, p = self.system.pupil(yo, l=wavelength, stop=stop)
        y, u = self.system.aim(yo, yp, z, p, filter=(not clip) if filter is None else filter)
        self.rays_given(y, u, wavelength, weight, ref)
        self.propagate(clip=clip)

    def rays_point(self, yo, wavelength=None, nrays=11, distribution="meridional", filter=None,
                   stop=None, clip=False):
        """a pupil distribution from one field point (geometric_trace.py:204-209)"""
        from rayopt.utils import pupil_distribution        # the reference's own grids
        ref, yp, weight = pupil_distribution(distribution, nrays)
        self.rays(yo, yp, wavelength, filter=filter, stop=stop, clip=clip, weight=weight, ref=ref)

    def rays_clipping(self, yo, wavelength=None, axis=1):
        """chief and the two rim rays along `axis` (geometric_trace.py:211-215)"""
        z, p = self.system.pupil(yo, l=wavelength, stop=-1)
        yp = np.zeros((3, 2))
        yp[1:, axis] = p[:, axis]/np.fabs(p).max()
        self.rays(yo, yp, wavelength, stop=-1, filter=False)

    def rays_line(self, yo, wavelength=None, nrays=21, eps=1e-2):
        """chief ray plus a meridional and a sagittal neighbour (pupil offset
        `eps`) for `nrays` field points from the axis to `yo`
        (geometric_trace.py:217-229); rows are grouped [chief | meridional |
        sagittal]"""
        s = self.system
        fields = np.linspace(0, 1, nrays)[:, None]*np.atleast_2d(yo)
        offsets = np.zeros((3, 2))
        offsets[1, 1] = offsets[2, 0] = eps
        y = np.empty((3, nrays, 3))
        u = np.empty((3, nrays, 3))
        z, p = s.pupil((0, 0), l=wavelength)
        reach = np.fabs(p).max()
        for k, f in enumerate(fields):
            z = s.aim_chief(f, z, reach, l=wavelength)
            y[:, k], u[:, k] = s.aim(f, offsets, z, p)
        self.rays_given(y.reshape(-1, 3), u.reshape(-1, 3), wavelength)
        self.propagate()

    def rays_paraxial(self, paraxial=None):
        """the two paraxial rays as real rays (geometric_trace.py:185-193)"""
        par = self.system.paraxial if paraxial is None else paraxial
        y = np.zeros((2, 2))
        u = np.zeros((2, 2))
        y[:, par.axis] = par.y[0]
        tan_u = np.asarray(par.u[0], float)
        u[:, par.axis] = tan_u*(1/np.sqrt(1 + np.square(tan_u)))   # sinarctan, utils.py:60-72
        self.rays_given(y, u)
        self.propagate()


def bind(reference_trace_class, engine=None, dtype=np.float64, exact=False, resident=False,
         alias_incidence=True):
    """Subclass of the reference's GeometricTrace with the hot path replaced.

        import rayopt, rayopt_b200
        GT = rayopt_b200.bind(rayopt.GeometricTrace)
        t = GT(system); t.rays_point((0, 1.), nrays=10**6, distribution="hexapolar")

    `resident=True`: the trace stays in HBM (``ResidentMixin``); ``t.y[-1]``,
    ``t.i[-1]``, ``t.rms()``, ``t.refocus()`` ... fetch or reduce single rows.
    """
    attrs = {"engine": engine, "exact": exact, "alias_incidence": alias_incidence,
             "__doc__": reference_trace_class.__doc__}
    if resident:
        if np.dtype(dtype) != np.float64:
            raise ValueError("the resident drop-in is FP64")
        return type("GeometricTrace", (ResidentMixin, reference_trace_class), attrs)
    attrs["dtype"] = dtype
    return type("GeometricTrace", (PropagateMixin, reference_trace_class), attrs)


def system_propagate(system, y, u, n, l, start=1, stop=None, clip=False,
                     engine=None, exact=False):
    """System.propagate (rayopt/system.py:459-464) as one launch: a generator
    yielding ``(y, u, n, i, t)`` per surface in the surface-normal frame.
    Input rays are in the lab frame of ``system[start-1]`` (i.e. AFTER its
    from_normal), as for the reference generator."""
    eng = engine or default_engine()
    table, ns, _ = pack_system(system, l, start, stop, n0=n)
    if len(table) == 0:
        return
    y, u = np.atleast_2d(y, u)
    Y, U, I, T = eng.trace(table, y, u, clip=clip, exact=exact)
    for j in range(len(table)):
        yield Y[j], U[j], ns[j], I[j], T[j]


def install(system_class, trace_class=None, engine=None, exact=False):
    """Monkey-patch a rayopt ``System`` class (and optionally its
    ``GeometricTrace``) so that every caller -- aim_chief / aim_marginal
    (rayopt/system.py:507-555), Analysis -- runs on the GPU engine."""
    def propagate(self, y, u, n, l, start=1, stop=None, clip=False):
        return system_propagate(self, y, u, n, l, start, stop, clip,
                                engine=engine, exact=exact)
    system_class.propagate = propagate
    if trace_class is not None:
        trace_class.allocate = PropagateMixin.allocate
        trace_class.propagate = PropagateMixin.propagate
        for k in ("_engine", "_empty", "_cache_system", "_materialize_i", "alias_incidence"):
            setattr(trace_class, k, getattr(PropagateMixin, k))
        trace_class.engine = engine
        trace_class.dtype = np.float64
        trace_class.exact = exact
