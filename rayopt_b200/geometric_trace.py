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
    """Standalone drop-in (no rayopt import needed): ``allocate / rays_given /
    propagate / rms / refocus`` on any `system` the packer can walk.  The ray
    launch helpers (``rays_point``, ``rays_clipping``, ``rays_line`` ...) are
    not restated here: with a rayopt ``System`` use ``bind(rayopt.
    GeometricTrace)``, which inherits the reference's own."""

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


def propagate_many(traces, clip=False):
    """``propagate(clip=clip)`` of several traces of ONE lens in one batched
    call (Engine.trace_bundles -> rtx_trace_batch_host): the front end for
    Analysis-style loops over fields x wavelengths of small bundles
    (rayopt/analysis.py:226-245, 266-280) -- launch rays loaded with
    ``rays_given`` (or by the reference's ``rays`` up to its propagate), then

        traces = [GT(system) for _ in bundles]
        for t, (yo, l) in zip(traces, bundles):   # aiming stays host Python
            z, p = system.pupil(yo, l=l); y, u = system.aim(yo, yp, z, p)
            t.rays_given(y, u, l, weight, ref)
        rayopt_b200.propagate_many(traces, clip=True)

    Every trace ends up exactly as after its own ``propagate``."""
    if not traces:
        return
    eng = traces[0]._engine()
    packs = []
    for t in traces:
        t._cache_system()
        table, n, rot0 = pack_system(t.system, t.l, 1, None, n0=t.n[0])
        if rot0 is not None:
            raise ValueError("propagate_many: a rotated object frame needs per-trace propagate")
        packs.append((table, n))
        if getattr(t, "_i_alias", False) and (table["flags"] & 1).any():
            t._materialize_i()
    res = eng.trace_bundles([p[0] for p in packs], [t.y[0] for t in traces],
                            [t.u[0] for t in traces], clip=clip, exact=traces[0].exact)
    for t, (table, n), (Y, U, I, T) in zip(traces, packs, res):
        rows = len(table)
        t.y[1:1 + rows], t.u[1:1 + rows], t.t[1:1 + rows] = Y, U, T
        if not getattr(t, "_i_alias", False):
            t.i[1:1 + rows] = I
        t.n[1:1 + rows] = n


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
