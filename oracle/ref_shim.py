"""Import shim for the LIVE reference (quartiq/rayopt).

TEST / BENCH INFRASTRUCTURE ONLY.  Used to (a) pin the oracle restatements in
this directory against the reference itself, (b) generate the golden fixtures
under tests/golden/, (c) run the reference's own GeometricTrace / System next
to (and bound to) the CUDA engine in the ``-m gpu`` tests and (d) time the
reference's CPU path in bench.py (``cpu_baseline`` / ``--impl reference``).
It is never imported by the product (rayopt_b200/).

Where the reference comes from: /root/reference in the build container; on the
GPU box (no /root/reference) the byte-for-byte copy staged by
oracle/make_ref.py under the git-ignored oracle/_ref/.

Recipe: stub `fastcache`, register a synthetic package
`rayopt` whose __path__ is the reference tree so that rayopt/__init__.py (which
pulls matplotlib / sqlalchemy / cython) is bypassed, never write bytecode into
the read-only tree.
"""
import functools
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))


def _root():
    env = os.environ.get("RAYOPT_REFERENCE")
    for cand in (env, "/root/reference", os.path.join(HERE, "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "rayopt", "geometric_trace.py")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _root()


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "rayopt", "geometric_trace.py"))


def load():
    """Return a namespace with the reference's hot-path classes."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    if "fastcache" not in sys.modules:
        fc = types.ModuleType("fastcache")
        fc.clru_cache = functools.lru_cache
        sys.modules["fastcache"] = fc
    if "rayopt" not in sys.modules:
        pkg = types.ModuleType("rayopt")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "rayopt")]
        sys.modules["rayopt"] = pkg
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from rayopt.system import System
        from rayopt.geometric_trace import GeometricTrace
        from rayopt.elements import Spheroid, Element, Interface
        from rayopt.material import Material
        from rayopt import utils, conjugates, pupils
    ns = types.SimpleNamespace(
        System=System, GeometricTrace=GeometricTrace, Spheroid=Spheroid,
        Element=Element, Interface=Interface, Material=Material,
        utils=utils, conjugates=conjugates, pupils=pupils)
    return ns
