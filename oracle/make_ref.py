#!/usr/bin/env python
"""Stage the UNMODIFIED reference (quartiq/rayopt) under oracle/_ref/ so that it
travels to the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY.  The reference is pure Python (the hot path
is numpy + one SciPy call), so "building" it is a file copy: the package's
``*.py`` modules are copied byte for byte from /root/reference/rayopt (read
only, present in the build container only) into the git-ignored
``oracle/_ref/rayopt/``.  Nothing under oracle/_ref/ is committed; it is a
build artefact like librtx.so and ships with the gpurun snapshot.

    python oracle/make_ref.py            # (re)stage when the tree is present

Users: oracle/ref_shim.py (import recipe of SURVEY.md Appendix B) for
* bench.py --impl reference / cpu_baseline  -> kind "reference",
* the ``-m gpu`` tests that bind the reference's own GeometricTrace / System
  to the CUDA engine and compare with the reference run on the host.

Not staged: the vendored C extension ``_transformations.c`` (the reference's
pure-Python fallback is used), ``library.sqlite`` (glass
catalog; the fixtures use Abbe-model glasses), the Cython simplex helper and
the reference's test-suite.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(os.environ.get("RAYOPT_REFERENCE", "/root/reference"), "rayopt")
DST = os.path.join(HERE, "_ref", "rayopt")


def stage(verbose=False):
    """copy SRC/*.py -> DST; returns the number of modules staged, or -1 when
    the reference tree is absent (GPU box: the prebuilt copy is used)"""
    if not os.path.isdir(SRC):
        return -1
    os.makedirs(DST, exist_ok=True)
    count = 0
    for name in sorted(os.listdir(SRC)):
        if not name.endswith(".py"):
            continue
        src, dst = os.path.join(SRC, name), os.path.join(DST, name)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
        count += 1
    for name in os.listdir(DST):                 # drop modules that vanished upstream
        if name.endswith(".py") and not os.path.exists(os.path.join(SRC, name)):
            os.remove(os.path.join(DST, name))
    if verbose:
        print("staged %d reference modules into %s" % (count, DST))
    return count


if __name__ == "__main__":
    n = stage(verbose=True)
    if n < 0:
        print("reference tree not present at %s (nothing staged)" % SRC)
    sys.exit(0)
