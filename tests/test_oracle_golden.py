"""The oracle (oracle/np_oracle.py) pinned against the committed golden
vectors, which tests/golden/make_golden.py generated from the LIVE reference
(quartiq/rayopt).  CPU only."""
import numpy as np
import pytest

import np_oracle
from conftest import golden_names, load_golden, assert_parity


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(name):
    c = load_golden(name)
    Y, U, I, T = np_oracle.trace(c["table"], c["y0"], c["u0"], clip=c["clip"],
                                 rot0=c["rot0"])
    newton = bool((c["table"]["n_asph"] >= 0).any())
    if not c["rotated"] and not newton:
        # same numpy expressions in the same order: bit-identical
        for a, b, w in ((Y, c["Y"], "y"), (U, c["U"], "u"), (I, c["I"], "i"),
                        (T, c["T"], "t")):
            assert np.array_equal(a, b, equal_nan=True), "%s %s not bit-exact" % (name, w)
    else:
        # np.dot (BLAS) may order / fuse the 3-term sums differently from the
        # restatement (rotations, and fprime in the Newton intercept,
        # elements.py:342): a few ulp, never more
        for a, b, w in ((Y, c["Y"], "y"), (U, c["U"], "u"), (I, c["I"], "i"),
                        (T, c["T"], "t")):
            assert_parity(a, b, 1e-13, "%s %s" % (name, w))


def test_known_answer_rms():
    """rayopt/test/test_raytrace.py:192-195: rms of the 13-ray radau bundle at
    field (0, 1) of the Cooke triplet is 0.052 +- 1 %."""
    c = load_golden("cooke_radau13")
    Y, U, I, T = np_oracle.trace(c["table"], c["y0"], c["u0"], clip=False)
    rms = np_oracle.rms(Y[-1], c["w"], ref=None)
    assert abs(rms - 0.052)/0.052 < 1e-2
    assert abs(rms - c["meta"]["rms"]) < 1e-15
    # ...and the 500-ray square grid agrees with it to 5 % (test_raytrace.py:196-199)
    q = load_golden("cooke_square500")
    Yq = np_oracle.trace(q["table"], q["y0"], q["u0"], clip=False)[0]
    assert abs(np_oracle.rms(Yq[-1], q["w"]) - rms)/rms < 5e-2


def test_golden_covers_edge_cases():
    """the fixtures exercise every branch of the path"""
    seen = set()
    for name in golden_names():
        c = load_golden(name)
        t = c["table"]
        if (t["n_asph"] >= 0).any():
            seen.add("newton")
        if ((t["k"] != 0) & (t["n_asph"] < 0)).any():
            seen.add("conic")
        if (t["flags"] & 1).any():
            seen.add("rotated")
        if (t["flags"] & 2).any():
            seen.add("alt")
        if (t["mu"] == -1).any():
            seen.add("mirror")
        if ((t["c"] == 0) & (t["mu"] != 1)).any():
            seen.add("plane_refract")
        if c["rot0"] is not None:
            seen.add("rot0")
        if c["clip"] and np.isnan(c["U"]).any():
            seen.add("vignette")
        if c["y0"].shape[0] == 1:
            seen.add("single")
    assert seen >= {"newton", "conic", "rotated", "alt", "mirror",
                    "plane_refract", "rot0", "vignette", "single"}, seen
