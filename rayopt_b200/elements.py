"""Element-level entry points of the reference on the GPU engine.

The third level of the drop-in boundary: callers that talk to a
single surface directly --

* ``InfiniteConjugate.aim`` -> ``surface.intercept(y, u)``  (rayopt/conjugates.py:254)
* ``GeometricTrace.opd``    -> ``Spheroid(curvature=1/radius).intercept`` (rayopt/geometric_trace.py:124)
* ``test_elements.py:129``  -> ``Spheroid.propagate(y0, u0, n0, l, clip)``

Each call is a one-surface trace (a table of one record in the element's own
normal frame: no offset, no rotation), i.e. exactly ``Interface.propagate``
(rayopt/elements.py:306-315) for that element.
"""
import numpy as np

from .engine import default_engine
from .surface_table import SURFACE_DTYPE, pack_element


class _Bare:
    """view of an element with the frame stripped (Element.propagate works in
    the element's normal frame: the caller has already applied to_normal)"""
    offset = (0., 0., 0.)
    rotated = False

    def __init__(self, e, mu=None):
        self._e = e
        if mu is not None:
            # explicit mu (Element.intercept / refract signatures carry no
            # wavelength): the material must not be evaluated -- a dispersive
            # one would need `l` (the reference's intercept(y, u) needs none,
            # conjugates.py:254)
            self.get_n_mu = lambda n0, l: (n0, mu)

    def __getattr__(self, k):
        return getattr(self._e, k)


def _record(element, n0, l, mu=None):
    t = np.zeros(1, SURFACE_DTYPE)
    n = pack_element(t[0], _Bare(element, mu), n0, l)
    return t, n


def propagate(element, y0, u0, n0, l, clip=True, engine=None, exact=False):
    """Interface.propagate / Element.propagate (rayopt/elements.py:230-236,
    306-315): returns ``(y, u, n, t*n0)`` in the element's normal frame."""
    eng = engine or default_engine()
    t, n = _record(element, n0, l)
    Y, U, I, T = eng.trace(t, np.atleast_2d(y0), np.atleast_2d(u0), clip=clip, exact=exact,
                           want=("y", "u", "t"))
    return Y[0], U[0], n, T[0]


def intercept(element, y, u, engine=None, exact=False):
    """Element.intercept / Spheroid.intercept / Interface.intercept
    (rayopt/elements.py:195-201, 477-501, 333-349): ray length to the surface."""
    eng = engine or default_engine()
    t, _ = _record(element, 1., None, mu=1.)
    t["n0"] = 1.
    Y, U, I, T = eng.trace(t, np.atleast_2d(y), np.atleast_2d(u), clip=False, exact=exact,
                           want=("t",))
    return T[0]


def refract(element, y, u0, mu, engine=None, exact=False):
    """Interface.refract (rayopt/elements.py:351-369) for points `y` ON the
    surface (the engine re-intercepts, which moves an on-surface point by
    ~1 ulp; the reference evaluates the normal at `y` as given)."""
    if mu == 1:
        return u0
    eng = engine or default_engine()
    t, _ = _record(element, 1., None, mu=mu)
    Y, U, I, T = eng.trace(t, np.atleast_2d(y), np.atleast_2d(u0), clip=False, exact=exact,
                           want=("u",))
    return U[0]
