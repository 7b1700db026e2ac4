"""Pins the ray-generator restatement (oracle/aim_oracle.py evaluating the
`rtx_aim` records of rayopt_b200/rays.py the way the CUDA kernels do) against
the LIVE reference: ``pupil_distribution`` (rayopt/utils.py:118-199),
``Pupil.map`` with and without its filter (rayopt/pupils.py:97-107),
``InfiniteConjugate.aim`` in all five projections with plane and curved object
surfaces, ``FiniteConjugate.aim`` with regular and telecentric pupils
(rayopt/conjugates.py:137-166, 208-255).  Bit-exact where no transcendental
function is evaluated per ray."""
import warnings

import numpy as np
import pytest

import aim_oracle
import ref_shim
from rayopt_b200.rays import aim_record, grid_spec

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")

P = np.array(((-3., -2.5), (2., 2.8)))          # pupil half-apertures [[-sag,-mer],[+sag,+mer]]


@pytest.fixture(scope="module")
def R():
    warnings.simplefilter("ignore")
    np.seterr(all="ignore")
    return ref_shim.load()


@pytest.mark.parametrize("dist,n", [
    ("half-meridional", 7), ("meridional", 12), ("sagittal", 9), ("cross", 23), ("tee", 152),
    ("square", 500), ("triangular", 700), ("hexapolar", 400), ("meridional", 1)])
def test_grids_equal_pupil_distribution(R, dist, n):
    """the per-candidate formulas (k*step + start linspace / mgrid points, the
    unit-circle predicate, prepended centre ray) reproduce the reference grid
    bit for bit, in order, with the same `ref` index"""
    ref, xy, weight = R.utils.pupil_distribution(dist, n)
    r, grid = grid_spec(dist, n)
    assert r == ref and weight is None
    if grid is None:
        assert n == 1
        return
    rec = aim_record(R.conjugates.InfiniteConjugate(angle=.2), (0, .5), 30., P, grid)
    px, py, keep = aim_oracle.candidates(rec[0])
    got = np.c_[px[keep], py[keep]]
    assert got.shape == xy.shape
    if dist == "hexapolar":                       # sin/cos of the ring angles: last ulps
        np.testing.assert_allclose(got, xy, rtol=0, atol=3e-16)
    else:
        assert np.array_equal(got, xy)


def test_quadrature_distributions_stay_on_the_host(R):
    for d in ("radau", "lobatto"):
        assert grid_spec(d, 13) == (0, None)


@pytest.mark.parametrize("filt", [False, True])
@pytest.mark.parametrize("projection", ["rectilinear", "stereographic", "equisolid",
                                        "orthographic", "equidistant"])
def test_infinite_conjugate_all_projections(R, projection, filt):
    obj = R.conjugates.InfiniteConjugate(angle=.35, projection=projection)
    plane = R.Spheroid()
    if projection == "orthographic":
        # the reference itself cannot aim this projection: conjugates.py:225-226
        # stacks a (n,1,1) array (``np.sqrt(1 - r)[:, None]`` with r already
        # (n,1)) and raises.  rays.project restates the evident intent
        # u = (y, sqrt(1 - |y|^2)); checked here for what it must satisfy.
        with pytest.raises(ValueError):
            obj.aim((0, .7), np.zeros((3, 2)), 25., P, surface=plane, filter=filt)
        from rayopt_b200.rays import project
        u = project(np.array([[0, .7], [-.4, .9]]), .35, projection)
        np.testing.assert_allclose(np.square(u).sum(1), 1, rtol=0, atol=1e-15)
        np.testing.assert_allclose(u[:, :2], np.array([[0, .7], [-.4, .9]])*np.sin(.35))
        return
    for dist, n in (("square", 300), ("tee", 31), ("hexapolar", 200)):
        ref, xy, _ = R.utils.pupil_distribution(dist, n)
        for yo in ((0., 0.), (0, .7), (-.4, .9)):
            want_y, want_u = obj.aim(yo, xy, 25., P, surface=plane, filter=filt)
            rec = aim_record(obj, yo, 25., P, grid_spec(dist, n)[1], filt, plane)
            y, u, pupil = aim_oracle.generate(rec)
            assert y.shape == want_y.shape, (dist, yo, y.shape, want_y.shape)
            if dist == "hexapolar":
                np.testing.assert_allclose(y, want_y, rtol=0, atol=1e-14, equal_nan=True)
            else:
                assert np.array_equal(y, want_y, equal_nan=True), (projection, dist, yo)
            assert np.array_equal(u, want_u)


@pytest.mark.parametrize("kw", [dict(curvature=.02), dict(curvature=-.03, conic=-.6),
                                dict(curvature=.01, aspherics=[0, 2e-6, -1e-9])])
def test_infinite_conjugate_curved_object_surface(R, kw):
    """y += surface.intercept(y, u) u (conjugates.py:254) with a sphere, a conic
    and an asphere as system[0]"""
    obj = R.conjugates.InfiniteConjugate(angle=.2)
    surf = R.Spheroid(**kw)
    ref, xy, _ = R.utils.pupil_distribution("square", 200)
    want_y, want_u = obj.aim((0, .8), xy, 20., P, surface=surf, filter=True)
    rec = aim_record(obj, (0, .8), 20., P, grid_spec("square", 200)[1], True, surf)
    assert rec["curved"][0] == 1
    y, u, _ = aim_oracle.generate(rec)
    if "aspherics" in kw:                        # Newton: the reference's fprime is a BLAS dot
        np.testing.assert_allclose(y, want_y, rtol=0, atol=1e-13)
    else:
        assert np.array_equal(y, want_y)
    assert np.array_equal(u, want_u)


@pytest.mark.parametrize("telecentric", [False, True])
@pytest.mark.parametrize("z", [40., -35.])
def test_finite_conjugate(R, z, telecentric):
    """FiniteConjugate.aim: object point, pupil angles arctan2(a, z), tan() of
    the mapped coordinates (last ulps), telecentric pupils, curved object
    surfaces (y_z = -surface_sag), z < 0"""
    obj = R.conjugates.FiniteConjugate(radius=6., pupil=dict(type="radius", radius=3.,
                                                             telecentric=telecentric))
    for surf in (R.Spheroid(), R.Spheroid(curvature=.02, conic=.3)):
        for dist, n, filt in (("square", 300, True), ("cross", 21, False), ("triangular", 150, False)):
            ref, xy, _ = R.utils.pupil_distribution(dist, n)
            want_y, want_u = obj.aim((.3, -.6), xy, z, P, surface=surf, filter=filt)
            rec = aim_record(obj, (.3, -.6), z, P, grid_spec(dist, n)[1], filt, surf)
            y, u, _ = aim_oracle.generate(rec)
            assert y.shape == want_y.shape
            assert np.array_equal(y, want_y)
            np.testing.assert_allclose(u, want_u, rtol=0, atol=3e-16)


def test_given_pupil_coordinates_and_random(R):
    obj = R.conjugates.InfiniteConjugate(angle=.3)
    rng = np.random.default_rng(4)
    yp = rng.uniform(-1, 1, (500, 2))
    want_y, want_u = obj.aim((0, .5), yp, 30., P, surface=R.Spheroid(), filter=True)
    rec = aim_record(obj, (0, .5), 30., P, None, True, R.Spheroid())
    y, u, pupil = aim_oracle.generate(rec, yp)
    assert 0 < len(y) < 500 and np.array_equal(y, want_y) and np.array_equal(u, want_u)
    # "random": uniform in the unit disc, centre ray first, reproducible from the seed
    rec = aim_record(obj, (0, .5), 30., P, grid_spec("random", 4000)[1], False, None, seed=7)
    _, _, p1 = aim_oracle.generate(rec)
    _, _, p2 = aim_oracle.generate(rec)
    assert np.array_equal(p1, p2) and p1.shape == (4001, 2) and np.all(p1[0] == 0)
    r2 = np.square(p1[1:]).sum(1)
    assert r2.max() <= 1 and abs(r2.mean() - .5) < .02 and abs(p1[1:].mean()) < .02
    rec2 = aim_record(obj, (0, .5), 30., P, grid_spec("random", 4000)[1], False, None, seed=8)
    assert not np.array_equal(aim_oracle.generate(rec2)[2], p1)
