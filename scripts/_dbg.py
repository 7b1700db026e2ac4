import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "oracle")
import numpy as np, np_oracle, test_gpu_random_systems as T
from rayopt_b200.engine import Engine
eng = Engine(0)
for seed, general in ((1, False), (8, True), (15, True)):
    if not general:
        rng = np.random.default_rng(1000 + seed); S = int(rng.integers(2, 24)); table = T.random_table(rng, S, False, False)
        n = int(rng.choice([257, 2000, 40003])); y0, u0 = T.random_rays(rng, n); rot0 = None
    else:
        rng = np.random.default_rng(2000 + seed); S = int(rng.integers(2, 16)); table = T.random_table(rng, S, True, True)
        rot0 = T.euler(*rng.normal(0, .02, 3)) if seed % 4 == 0 else None
        n = int(rng.choice([300, 5000])); y0, u0 = T.random_rays(rng, n)
    clip = bool(seed % 2)
    want = np_oracle.trace(table, y0, u0, clip=clip, rot0=rot0)
    ok = T.well_conditioned(table, y0, u0, want, clip, rot0, amp=30 if general else 2e3)
    got = eng.trace(table, y0, u0, clip=clip, rot0=rot0, dtype=np.float32 if general else np.float64)
    for a, b, w in zip(got, want, "yuit"):
        a = a.astype(float).reshape(a.shape[0], a.shape[1], -1); b = b.reshape(a.shape)
        scale = np.maximum(np.nanmax(np.where(np.isfinite(b), np.abs(b), 0), axis=(1, 2), keepdims=True), 1.0)
        if general: scale = np.full_like(scale, scale.max())
        d = np.nan_to_num(np.abs(a - b)/np.maximum(np.abs(b), scale))
        d[:, ~ok] = 0
        j, r, c = np.unravel_index(np.argmax(d), d.shape)
        print("seed", seed, w, "worst %.3e at surface %d ray %d comp %d  got %r want %r scale %.3g" % (d[j, r, c], j, r, c, a[j, r, c], b[j, r, c], scale[j, 0, 0]))
        if d[j, r, c] > (1e-5 if general else 1e-10):
            rec = table[j]
            print("   surface: c=%g k=%g n_asph=%d flags=%d mu=%g off=%s" % (rec["c"], rec["k"], rec["n_asph"], rec["flags"], rec["mu"], rec["offset"]))
            for jj in range(max(0, j - 2), j + 1):
                print("   s%d: y=%s u=%s t=%r | want y=%s t=%r kind c=%g k=%g asph=%d" % (jj, got[0][jj, r], got[1][jj, r], got[3][jj, r], want[0][jj, r], want[3][jj, r], table[jj]["c"], table[jj]["k"], table[jj]["n_asph"]))
