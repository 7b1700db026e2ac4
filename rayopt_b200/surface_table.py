"""System -> rtx_surface[] packer (host side of the drop-in boundary).

Walks a rayopt ``System`` (or anything that quacks like one: a sequence of
elements with the attributes listed below) exactly the way
``System.propagate`` (rayopt/system.py:459-464) and ``Interface.propagate``
(rayopt/elements.py:306-315) do, and emits one POD record per traced surface
for ONE wavelength.  ``System`` is mutable shared state (``refocus`` edits
``system[at].distance``, rayopt/geometric_trace.py:98), so the table is
re-packed on every ``propagate`` call; it is S small records.

The "derived" members are computed here with the SAME Python expressions the
reference evaluates per call (e.g. ``(1 + k)*c**2``, ``mu**2 - 1``), so that
the engine's RTX_EXACT mode can be bit-identical to the reference.
"""
import numpy as np

RTX_MAX_ASPH = 10
RTX_MAX_SURFACES = 256
F_ROTATED = 1
F_ALT = 2

# mirrors `struct rtx_surface` in include/rtx.h (natural C alignment; the size
# is cross-checked against rtx_sizeof_surface() when the library is loaded)
SURFACE_DTYPE = np.dtype([
    ("offset", "<f8", (3,)),
    ("rot", "<f8", (9,)),
    ("c", "<f8"),
    ("k", "<f8"),
    ("kc2", "<f8"),
    ("radius2", "<f8"),
    ("mu", "<f8"),
    ("muf", "<f8"),
    ("sgn", "<f8"),
    ("mu2m1", "<f8"),
    ("n0", "<f8"),
    ("n", "<f8"),
    ("asph", "<f8", (RTX_MAX_ASPH,)),
    ("dasph", "<f8", (RTX_MAX_ASPH,)),
    ("n_asph", "<i4"),
    ("flags", "<u4"),
], align=True)


def get_n_mu(element, n0, l):
    """Interface.get_n_mu (rayopt/elements.py:283-289), duck-typed; a plain
    ``Element`` has no material and never refracts (elements.py:230-236)."""
    fn = getattr(element, "get_n_mu", None)
    if fn is not None:
        n, mu = fn(n0, l)
        return n, mu
    return n0, 1.


def pack_element(rec, e, n0, l):
    """Fill one record from element `e` hit from a medium of index `n0` at
    wavelength `l`; returns the index after the surface."""
    rec["offset"] = np.asarray(e.offset, float)            # system.py:461
    flags = 0
    if getattr(e, "rotated", False):                       # elements.py:135
        flags |= F_ROTATED
        rec["rot"] = np.asarray(e.rot_normal, float).reshape(9)
    else:
        rec["rot"] = np.eye(3).reshape(9)
    c = getattr(e, "curvature", 0.)
    k = getattr(e, "conic", 0.)
    asph = getattr(e, "aspherics", None)
    if getattr(e, "alternate_intersection", False):        # elements.py:497
        flags |= F_ALT
    rec["c"] = c
    rec["k"] = k
    rec["kc2"] = (1 + k)*c**2                              # elements.py:448,467
    radius = getattr(e, "radius", np.inf)
    rec["radius2"] = radius**2                             # elements.py:207
    n, mu = get_n_mu(e, n0, l)
    if not mu:                                             # elements.py:313
        mu = 1.
    rec["mu"] = mu
    rec["muf"] = abs(mu)                                   # elements.py:360
    rec["sgn"] = np.sign(mu)                               # elements.py:367
    rec["mu2m1"] = mu**2 - 1                               # elements.py:366
    rec["n0"] = n0
    rec["n"] = n
    rec["asph"] = 0.
    rec["dasph"] = 0.
    if asph is None:
        rec["n_asph"] = -1                                 # elements.py:478
    else:
        asph = list(asph)
        if len(asph) > RTX_MAX_ASPH:
            raise ValueError("at most %d aspheric coefficients are supported, "
                             "got %d" % (RTX_MAX_ASPH, len(asph)))
        rec["n_asph"] = len(asph)
        for i, a in enumerate(asph):
            rec["asph"][i] = a
            rec["dasph"][i] = 2*(i + 1)*a                  # elements.py:472
    rec["flags"] = flags
    return n


def pack_system(system, l, start=1, stop=None, n0=None):
    """Records for ``system[start:stop]`` at wavelength `l`.

    `n0` is the index in front of ``system[start]``; default
    ``system.refractive_index(l, start - 1)`` walked the way
    ``GeometricTrace.rays_given`` does for start == 1
    (rayopt/geometric_trace.py:69).

    Returns ``(table, n, rot0)``: the structured array, the (S,) array of
    indices after each surface (what the reference stores into
    ``GeometricTrace.n[start:stop]``) and the 3x3 ``rot_normal`` of
    ``system[start-1]`` or None (geometric_trace.py:76).
    """
    if hasattr(system, "pack"):                # pre-packed (PackedSystem)
        return system.pack(l, start, stop, n0)
    elements = list(system[start:stop])
    S = len(elements)
    if S > RTX_MAX_SURFACES:
        raise ValueError("too many surfaces: %d" % S)
    if n0 is None:
        n0 = system.refractive_index(l, start - 1)
    # column-wise fill (the record-by-record pack_element is ~13 us per
    # surface in numpy scalar assignments; aiming issues hundreds of tiny
    # traces, rayopt/system.py:507-555).  Same expressions as pack_element.
    table = np.zeros(S, SURFACE_DTYPE)
    n = np.empty(S)
    off, rot, cs, ks, kc2, rad2, mus, n0s, flags, nas = [], [], [], [], [], [], [], [], [], []
    eye = (1., 0., 0., 0., 1., 0., 0., 0., 1.)
    for j, e in enumerate(elements):
        off.append(e.offset)
        f = 0
        if getattr(e, "rotated", False):
            f |= F_ROTATED
            rot.append(np.asarray(e.rot_normal, float).reshape(9))
        else:
            rot.append(eye)
        c = getattr(e, "curvature", 0.)
        k = getattr(e, "conic", 0.)
        if getattr(e, "alternate_intersection", False):
            f |= F_ALT
        cs.append(c)
        ks.append(k)
        kc2.append((1 + k)*c**2)                           # elements.py:448,467
        radius = getattr(e, "radius", np.inf)
        rad2.append(radius**2)                             # elements.py:207
        nn, mu = get_n_mu(e, n0, l)
        if not mu:                                         # elements.py:313
            mu = 1.
        mus.append(mu)
        n0s.append(n0)
        n[j] = n0 = nn
        flags.append(f)
        asph = getattr(e, "aspherics", None)
        if asph is None:
            nas.append(-1)
        else:
            asph = list(asph)
            if len(asph) > RTX_MAX_ASPH:
                raise ValueError("at most %d aspheric coefficients are supported, "
                                 "got %d" % (RTX_MAX_ASPH, len(asph)))
            nas.append(len(asph))
            for i, a in enumerate(asph):
                table["asph"][j, i] = a
                table["dasph"][j, i] = 2*(i + 1)*a         # elements.py:472
    if S:
        table["offset"] = off
        table["rot"] = rot
        table["c"], table["k"], table["kc2"], table["radius2"] = cs, ks, kc2, rad2
        table["mu"] = mus
        table["muf"] = [abs(m) for m in mus]               # elements.py:360
        table["sgn"] = [np.sign(m) for m in mus]           # elements.py:367
        table["mu2m1"] = [m**2 - 1 for m in mus]           # elements.py:366
        table["n0"], table["n"] = n0s, n
        table["n_asph"], table["flags"] = nas, flags
    init = system[start - 1]
    rot0 = None
    if getattr(init, "rotated", False):
        rot0 = np.ascontiguousarray(init.rot_normal, float)
    return table, n, rot0


class PackedSystem:
    """A lens given directly as packed tables, one per wavelength -- what
    ``pack_system`` would produce from a rayopt ``System``.  Lets
    ``GeometricTrace`` run where rayopt itself is not installed (the GPU
    benchmark box) and skips the per-call walk over the elements."""

    def __init__(self, wavelengths, tables, n_before_first):
        self.wavelengths = [float(l) for l in wavelengths]
        self._tables = {float(l): np.ascontiguousarray(t, SURFACE_DTYPE)
                        for l, t in zip(wavelengths, tables)}
        self._n0 = {float(l): float(n) for l, n in zip(wavelengths, n_before_first)}
        self._len = len(tables[0]) + 1

    def __len__(self):
        return self._len

    def refractive_index(self, l, i):
        t = self._tables[float(l)]
        return self._n0[float(l)] if i == 0 else float(t["n"][i - 1])

    def pack(self, l, start=1, stop=None, n0=None):
        full = self._tables[float(l)]
        idx = range(self._len)[start:stop]
        table = full[idx.start - 1:idx.stop - 1].copy() if len(idx) else full[:0].copy()
        rot0 = None
        if start >= 2 and full["flags"][start - 2] & F_ROTATED:
            rot0 = full["rot"][start - 2].reshape(3, 3).copy()
        return table, table["n"].copy(), rot0


def table_to_json(table):
    """Lossless (repr-exact floats) JSON-able form of a table."""
    out = []
    for rec in table:
        d = {}
        for name in SURFACE_DTYPE.names:
            v = rec[name]
            if isinstance(v, np.ndarray):
                d[name] = [float(x).hex() for x in v]
            elif name in ("n_asph", "flags"):
                d[name] = int(v)
            else:
                d[name] = float(v).hex()
        out.append(d)
    return out


def table_from_json(items):
    table = np.zeros(len(items), SURFACE_DTYPE)
    for rec, d in zip(table, items):
        for name in SURFACE_DTYPE.names:
            v = d[name]
            if isinstance(v, list):
                rec[name] = [float.fromhex(x) for x in v]
            elif name in ("n_asph", "flags"):
                rec[name] = v
            else:
                rec[name] = float.fromhex(v)
    return table
