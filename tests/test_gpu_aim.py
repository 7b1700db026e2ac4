"""The device ray generator (rtx_aim_plan / rtx_aim_rays) against
its numpy restatement oracle/aim_oracle.py -- itself pinned bit for bit to the
reference's pupil_distribution / Pupil.map / Conjugate.aim in
tests/test_aim_oracle.py -- and, where the reference tree is available, against
the reference's own ``System.aim`` + ``rays_point``."""
import types
import warnings

import numpy as np
import pytest

import aim_oracle
import ref_shim
from rayopt_b200.rays import aim_record, grid_spec

pytestmark = pytest.mark.gpu

P = np.array(((-3., -2.5), (2., 2.8)))


@pytest.fixture(scope="module")
def eng():
    from rayopt_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def conj(finite=False, **kw):
    """duck-typed rayopt conjugate (what aim_record reads)"""
    pupil = types.SimpleNamespace(telecentric=kw.pop("telecentric", False))
    return types.SimpleNamespace(finite=finite, pupil=pupil, **kw)


def surf(**kw):
    d = dict(curvature=0., conic=0., aspherics=None, alternate_intersection=False, radius=np.inf)
    d.update(kw)
    return types.SimpleNamespace(**d)


def device(eng, rec, yp=None, dtype=np.float64, **kw):
    d_yp = None if yp is None else eng.to_device(yp)
    y, u, p = eng.aim_rays(rec, dtype, yp=d_yp, want_pupil=True, **kw)
    eng.sync()
    return y.download(), u.download(), p.download()


GRIDS = [("half-meridional", 7), ("meridional", 12), ("sagittal", 9), ("cross", 23),
         ("tee", 152), ("square", 5000), ("triangular", 7000), ("hexapolar", 4000)]


@pytest.mark.parametrize("filt", [False, True])
@pytest.mark.parametrize("dist,n", GRIDS)
def test_infinite_grids_bit_exact(eng, dist, n, filt):
    """every device grid, with and without the Pupil.map filter (ordered
    compaction), equals the restatement bit for bit: ray count, order, pupil
    coordinates, positions, directions"""
    obj = conj(angle=.35, projection="stereographic")
    rec = aim_record(obj, (-.4, .9), 25., P, grid_spec(dist, n)[1], filt, surf())
    wy, wu, wp = aim_oracle.generate(rec)
    assert eng.aim_count(rec) == len(wy)
    y, u, p = device(eng, rec)
    if dist == "hexapolar":                       # sincos of the ring angles: last ulps
        np.testing.assert_allclose(p, wp, rtol=0, atol=3e-16)
        np.testing.assert_allclose(y, wy, rtol=0, atol=1e-14)
    else:
        assert np.array_equal(p, wp) and np.array_equal(y, wy)
    assert np.array_equal(u, wu)


def test_large_mesh_compaction_and_subranges(eng):
    """a 3e6-candidate square mesh with the filter: the two-pass compaction
    keeps the reference order; any sub-range [first, first+count) can be
    generated on its own (ray sharding over GPUs)"""
    obj = conj(angle=.2, projection="rectilinear")
    n = 3_000_000
    rec = aim_record(obj, (0, .7), 30., P, grid_spec("square", n)[1], True, surf())
    wy, wu, wp = aim_oracle.generate(rec)
    total = eng.aim_count(rec)
    assert total == len(wy) and .5*n < total < n
    y, u, p = device(eng, rec)
    assert np.array_equal(y, wy) and np.array_equal(u, wu) and np.array_equal(p, wp)
    for first, count in ((0, 1), (1, 1023), (777_777, 100_001), (total - 5, 5)):
        ys, us, ps = device(eng, rec, first=first, count=count)
        assert np.array_equal(ys, wy[first:first + count]) and np.array_equal(ps, wp[first:first + count])
    y32, u32, _ = device(eng, rec, dtype=np.float32, first=1000, count=4096)
    assert y32.dtype == np.float32 and np.array_equal(y32, wy[1000:5096].astype(np.float32))


@pytest.mark.parametrize("kw", [dict(curvature=.02), dict(curvature=-.03, conic=-.6),
                                dict(curvature=.01, aspherics=[0, 2e-6, -1e-9])])
def test_curved_object_surface(eng, kw):
    """infinite object, curved system[0]: the rays are intercepted by the trace
    kernel's own surface_step (sphere / conic bit-exact, Newton to 1e-13)"""
    rec = aim_record(conj(angle=.2, projection="rectilinear"), (0, .8), 20., P,
                     grid_spec("square", 3000)[1], True, surf(**kw))
    assert rec["curved"][0] == 1
    wy, wu, wp = aim_oracle.generate(rec)
    y, u, p = device(eng, rec)
    if "aspherics" in kw:
        np.testing.assert_allclose(y, wy, rtol=0, atol=1e-13)
    else:
        assert np.array_equal(y, wy)
    assert np.array_equal(u, wu) and np.array_equal(p, wp)


@pytest.mark.parametrize("telecentric", [False, True])
@pytest.mark.parametrize("z", [40., -35.])
def test_finite_conjugate(eng, z, telecentric):
    obj = conj(True, radius=6., telecentric=telecentric)
    for dist, n, filt in (("square", 3000, True), ("cross", 21, False), ("triangular", 1500, False)):
        rec = aim_record(obj, (.3, -.6), z, P, grid_spec(dist, n)[1], filt, surf())
        wy, wu, wp = aim_oracle.generate(rec)
        y, u, p = device(eng, rec)
        assert np.array_equal(y, wy) and np.array_equal(p, wp)
        np.testing.assert_allclose(u, wu, rtol=0, atol=4e-16)        # tan(): last ulps
        np.testing.assert_allclose(np.square(u).sum(1), 1, rtol=0, atol=7e-16)


def test_given_coordinates_and_random(eng):
    obj = conj(angle=.3, projection="equisolid")
    yp = np.random.default_rng(4).uniform(-1, 1, (70001, 2))
    rec = aim_record(obj, (0, .5), 30., P, None, True, surf())
    wy, wu, wp = aim_oracle.generate(rec, yp)
    y, u, p = device(eng, rec, yp)
    assert 0 < len(y) < len(yp)
    assert np.array_equal(y, wy) and np.array_equal(u, wu) and np.array_equal(p, wp)
    # random: the counter-based generator is reproducible, uniform over the disc,
    # and its coordinates can be fed to the reference (they are returned)
    rec = aim_record(obj, (0, .5), 30., P, grid_spec("random", 200000)[1], False, surf(), seed=7)
    y, u, p = device(eng, rec)
    y2, u2, p2 = device(eng, rec)
    assert np.array_equal(p, p2) and p.shape == (200001, 2) and np.all(p[0] == 0)
    wy, wu, wp = aim_oracle.generate(rec)                           # same integer hash on the host
    np.testing.assert_allclose(p, wp, rtol=0, atol=1e-15)         # sincospi vs cos/sin(2 pi phi)
    r2 = np.square(p[1:]).sum(1)
    assert r2.max() <= 1 + 1e-15 and abs(r2.mean() - .5) < 2e-3 and np.abs(p[1:].mean(0)).max() < 3e-3
    h = np.histogram2d(p[1:, 0], p[1:, 1], bins=8, range=((-.7, .7), (-.7, .7)))[0]
    assert h.std()/h.mean() < .05                                   # flat inside the disc
    rec8 = aim_record(obj, (0, .5), 30., P, grid_spec("random", 200000)[1], False, surf(), seed=8)
    assert not np.array_equal(device(eng, rec8)[2], p)


@pytest.mark.skipif(not ref_shim.available(), reason="no reference tree")
def test_resident_rays_point_on_device_vs_reference(eng):
    """ResidentTrace.rays_point on the reference's own Systems: launch rays
    generated in HBM for every grid distribution (filter on and off), traced,
    and compared with the reference's rays_point of the same call"""
    import yaml
    import systems_yaml
    from conftest import assert_parity
    from rayopt_b200 import ResidentTrace
    warnings.simplefilter("ignore")
    np.seterr(all="ignore")
    R = ref_shim.load()
    for name in ("double_gauss", "cooke"):
        s = R.System(**yaml.safe_load(systems_yaml.SYSTEMS[name]))
        s.update()
        s.paraxial.refocus()
        l0 = eng.launch_count()
        for dist, n, clip, filt in (("hexapolar", 3000, True, None), ("square", 2000, False, None),
                                    ("triangular", 1500, True, True), ("tee", 152, True, None),
                                    ("cross", 21, False, False), ("meridional", 11, False, None)):
            ref = R.GeometricTrace(s)
            got = ResidentTrace(s, engine=eng, exact=True)
            ref.rays_point((0, .7), s.wavelengths[1], nrays=n, distribution=dist, clip=clip, filter=filt)
            got.rays_point((0, .7), s.wavelengths[1], nrays=n, distribution=dist, clip=clip, filter=filt)
            assert got.nrays == ref.y.shape[1] and got.ref == ref.ref, (name, dist)
            tol = 1e-11 if dist == "hexapolar" else 0
            for k in "yuit":
                a, b = np.asarray(getattr(got, k)), getattr(ref, k)
                if tol:
                    assert_parity(a, b, tol, "%s %s %s" % (name, dist, k))
                else:
                    assert np.array_equal(a, b, equal_nan=True), (name, dist, k)
            got.free()
        assert eng.launch_count() - l0 >= 12
