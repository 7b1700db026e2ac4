#!/usr/bin/env python
"""three wavelength bundles (C2): 3 launches vs one batched launch"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from rayopt_b200.engine import Engine
eng = Engine(0)
ent = bench.load_system("double_gauss")
S, N = ent["S"], 10_000_000
ld = (N + 63)//64*64
dev = []
for li in range(3):
    y0, u0 = bench.make_rays(ent, li, N, li)
    dev.append((eng.to_device(y0), eng.to_device(u0), eng.empty((S, ld, 3)), eng.empty((S, ld, 3)),
                eng.empty((S, ld, 3)), eng.empty((S, ld))))
def sep():
    for li in range(3):
        d = dev[li]
        eng.trace_device(ent["tables"][li], d[0], d[1], d[2], d[3], d[4], d[5], N=N, ld=ld, clip=True)
def bat():
    eng.trace_device_batch(ent["tables"][:3], [d[0] for d in dev], [d[1] for d in dev], [d[2] for d in dev],
                           [d[3] for d in dev], [d[4] for d in dev], [d[5] for d in dev], Ns=[N]*3, ld=ld, clip=True)
for name, fn in (("3 launches", sep), ("1 batched launch", bat), ("3 launches", sep), ("1 batched launch", bat)):
    for _ in range(5): fn()
    eng.sync(); eng.timer_start()
    for _ in range(30): fn()
    ms = eng.timer_stop()/30
    print("%-18s %.3f ms/step  %.3e ray-surf/s  %.1f GB/s" % (name, ms, 3*N*S/ms*1e3, 3*N*(48+80*S)/ms/1e6), flush=True)
