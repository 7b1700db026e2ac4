"""Import shim for the LIVE reference (quartiq/rayopt at /root/reference).

TEST INFRASTRUCTURE ONLY.  Used in the build container (where /root/reference
exists) to (a) pin the oracle restatements in this directory against the
reference itself and (b) generate the golden fixtures under tests/golden/.
It is never imported by the product (rayopt_b200/) and it cannot work on the
GPU box (no /root/reference there).

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

REFERENCE_ROOT = os.environ.get("RAYOPT_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "rayopt"))


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
