#!/bin/bash
# compute-sanitizer passes over the store paths (small problem sizes), and the small-N latency
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import load_golden, load_systems
from rayopt_b200.engine import Engine
from rayopt_b200.rays import aim_infinite, disc
eng = Engine(0)
ent = load_systems()["zoom"]
table, aim = ent["tables"][0], ent["aim"][0][2]
y0, u0 = aim_infinite(aim["field"], disc(40001, 1), aim["z"], aim["p"], ent["object_angle"])
ref = eng.trace(table, y0, u0, clip=True, direct=True)
for kw in (dict(), dict(rpt=1), dict(rpt=2), dict(exact=True), dict(dtype=np.float32), dict(keep_last=True)):
    out = eng.trace(table, y0, u0, clip=True, **kw)
    print(kw, out[0].shape, flush=True)
ya, ua = y0[:39936], u0[:39936]                         # FP32 four-rays-per-thread kernels (ld % 128 == 0)
d32 = [eng.to_device(ya, np.float32), eng.to_device(ua, np.float32)]
o32 = [eng.empty((len(table), 39936, 3), np.float32) for _ in range(3)] + [eng.empty((len(table), 39936), np.float32)]
eng.trace_device(table, d32[0], d32[1], *o32, N=39936, ld=39936, clip=True)
ca = load_systems()["cooke_asph"]["tables"][0]          # Newton kernels: FP64 2x8 per-CTA, FP32 4x8 per-warp
for dt in (np.float64, np.float32):
    dd = [eng.to_device(ya, dt), eng.to_device(ua, dt)]
    oo = [eng.empty((len(ca), 39936, 3), dt) for _ in range(3)] + [eng.empty((len(ca), 39936), dt)]
    eng.trace_device(ca, dd[0], dd[1], *oo, N=39936, ld=39936, clip=True)
eng.sync()
print("fp32 rpt4 / newton configs ok", flush=True)
# side outputs, fused gather destinations, batched bundles
d_y0, d_u0 = eng.to_device(y0), eng.to_device(u0)
mask = eng.empty(((len(y0) + 31)//32,), np.uint32)
acc = eng.empty((len(y0),))
eng.trace_device(table, d_y0, d_u0, None, None, None, None, N=len(y0), clip=True, mask=mask, path_sum=acc)
bufs = [eng.empty((len(y0) + 256, 3)) for _ in range(4)]
eng.trace_gather(table, d_y0, d_u0, [b.ptr for b in bufs[:2]], 128, clip=True,
                 dst_i_ptrs=[b.ptr for b in bufs[2:]])          # ragged N: per-ray stores
eng.trace_gather(table, d_y0, d_u0, [b.ptr for b in bufs[:2]], 128, N=40000//64*64, clip=True,
                 dst_i_ptrs=[b.ptr for b in bufs[2:]])          # whole groups: bulk stores
bxy = [eng.empty((len(y0) + 256, 2)) for _ in range(2)]
eng.trace_gather(table, d_y0, d_u0, [b.ptr for b in bxy], 128, N=40000//64*64, clip=True, xy=True)
eng.trace_gather(table, d_y0, d_u0, [b.ptr for b in bxy], 128, clip=True, xy=True)
# fused epilogues, ray generator (compaction), batched host front end
m = eng.trace_reduce(table, d_y0, d_u0, clip=True, center=np.zeros(4))
A, P = eng.empty((len(y0),)), eng.empty((len(y0), 3))
eng.trace_opd(table[:-1], d_y0, d_u0, dict(y0_ref=y0[0], u0_ref=u0[0], n0=1., n_after=1., M=np.eye(3),
              d=np.zeros(3), radius=-80., infinite=True), A, P, clip=True)
import types
from rayopt_b200.rays import aim_record, grid_spec
obj = types.SimpleNamespace(finite=False, angle=.2, projection="rectilinear", pupil=types.SimpleNamespace(telecentric=False))
pp = np.array(((-3., -2.5), (2., 2.8)))
for dist, n, filt in (("square", 300000, True), ("triangular", 50000, False), ("random", 70000, True), ("tee", 152, False)):
    rec = aim_record(obj, (0, .7), 30., pp, grid_spec(dist, n)[1], filt, None)
    yy, uu, pq = eng.aim_rays(rec, want_pupil=True)
    eng.sync()
    print(dist, yy.shape, flush=True)
outs_b = eng.trace_bundles([table]*11, [y0[:150 + 7*k] for k in range(11)], [u0[:150 + 7*k] for k in range(11)], clip=True)
print("epilogues / generator / batch host ok", m[5], flush=True)
S, ld = len(table), (len(y0) + 63)//64*64
outs = [[eng.empty((S, ld, 3)) for _ in range(3)] + [eng.empty((S, ld))] for _ in range(2)]
eng.trace_device_batch([table, ent["tables"][1]], [d_y0, d_y0], [d_u0, d_u0], [o[0] for o in outs],
                       [o[1] for o in outs], [o[2] for o in outs], [o[3] for o in outs],
                       Ns=[len(y0), 33000], ld=ld, clip=True)
eng.sync()
print("side outputs / gather / batch ok", flush=True)
c = load_golden("cooke_asph_f07_clip")
eng.trace(c["table"], c["y0"], c["u0"], clip=True)
c = load_golden("tilted_clip1")
eng.trace(c["table"], c["y0"], c["u0"], clip=True, rot0=c["rot0"])
eng.close()
print("done")
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done" gpurun_out/sanitizer_$tool.txt | tail -3
done
python - <<'PY' > gpurun_out/latency.txt 2>&1
import sys, time, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from conftest import load_golden
from rayopt_b200.engine import Engine
eng = Engine(0)
c = load_golden("cooke_single_ray")
for n in (1, 3, 13, 1000):
    y = np.repeat(c["y0"], n, 0); u = np.repeat(c["u0"], n, 0)
    for _ in range(20): eng.trace(c["table"], y, u)
    t0 = time.perf_counter()
    for _ in range(500): eng.trace(c["table"], y, u)
    print("N=%5d rays, S=8: %.1f us per GeometricTrace-style host-buffer trace call" % (n, (time.perf_counter()-t0)/500*1e6))
PY
cat gpurun_out/latency.txt
free -g | head -2; nproc
