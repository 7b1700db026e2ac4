#!/usr/bin/env python
"""CPU timing of the reference's hot path on all host cores.

TEST / BENCH INFRASTRUCTURE ONLY: used by bench.py's ``cpu_baseline`` and
``--impl reference`` legs, run as a separate process

    python oracle/cpu_bench.py --system double_gauss --field 0 0.7 \
        --rays-total 10000000 --steps 20 --warmup 5

and prints one JSON line.  Nothing in the product imports it.

kind "reference": the UNMODIFIED reference (oracle/_ref staged by
oracle/make_ref.py, or /root/reference) -- every worker process builds the
reference's own ``System`` from the fixture YAML (tests/golden/systems_yaml.py),
aims its own bundle with ``System.pupil`` / ``System.aim`` and then, per timed
step and wavelength, runs ``GeometricTrace.rays_given`` + ``propagate(clip=True)``
(rayopt/geometric_trace.py:49-80) exactly as a user would.  The reference's
trace is single-threaded numpy; "all host cores" is ray-sharded
multiprocessing: a persistent set of worker processes, one
bundle each, released together per step; the step time is the wall time until
the slowest worker is done.

kind "port": the numpy restatement oracle/np_oracle.py on the packed tables of
tests/golden/systems.json -- the fallback when no reference tree is available.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _disc(n, seed):
    """uniform pupil coordinates in the unit disc"""
    import numpy as np
    rng = np.random.default_rng(seed)
    r, phi = np.sqrt(rng.random(n)), 2*np.pi*rng.random(n)
    return np.c_[r*np.cos(phi), r*np.sin(phi)]


class _RefWorker:
    """the reference itself"""
    kind = "reference"

    def __init__(self, system, field, n, seed, n_warm):
        import warnings
        import yaml
        import ref_shim
        import systems_yaml
        warnings.simplefilter("ignore")
        import numpy as np
        np.seterr(all="ignore")
        R = ref_shim.load()
        s = R.System(**yaml.safe_load(systems_yaml.SYSTEMS[system]))
        s.update()
        s.paraxial.refocus()
        self.s, self.S = s, len(s) - 1
        self.rays, self.small = [], []
        for l in s.wavelengths:
            z, p = s.pupil(field, l=l)
            self.rays.append((l,) + tuple(s.aim(field, _disc(n, seed), z, p, filter=False)))
            self.small.append((l,) + tuple(s.aim(field, _disc(n_warm, seed), z, p,
                                                 filter=False)))
        self.g = R.GeometricTrace(s)
        self.gw = R.GeometricTrace(s)

    def step(self, warm=False):
        g, rays = (self.gw, self.small) if warm else (self.g, self.rays)
        t0 = time.perf_counter()
        for l, y, u in rays:
            g.rays_given(y, u, l)
            g.propagate(clip=True)
        return time.perf_counter() - t0


class _PortWorker:
    """numpy restatement on the committed tables"""
    kind = "port"

    def __init__(self, system, field, n, seed, n_warm):
        import numpy as np
        from rayopt_b200.rays import aim_infinite
        from rayopt_b200.surface_table import table_from_json
        np.seterr(all="ignore")
        with open(os.path.join(ROOT, "tests", "golden", "systems.json")) as f:
            ent = json.load(f)[system]
        self.tables = [table_from_json(t) for t in ent["tables"]]
        self.S = ent["S"]
        fi = min(range(len(ent["aim"][0])),
                 key=lambda k: sum((a - b)**2 for a, b in zip(ent["aim"][0][k]["field"], field)))
        self.rays, self.small = [], []
        for li in range(len(self.tables)):
            aim = ent["aim"][li][fi]
            for dst, m in ((self.rays, n), (self.small, n_warm)):
                dst.append(aim_infinite(aim["field"], _disc(m, seed), aim["z"], aim["p"],
                                        ent["object_angle"]))

    def step(self, warm=False):
        import np_oracle
        t0 = time.perf_counter()
        for table, (y, u) in zip(self.tables, self.small if warm else self.rays):
            np_oracle.trace(table, y, u, clip=True)
        return time.perf_counter() - t0


def _worker(conn, kind, system, field, n, seed, n_warm):
    try:
        w = (_RefWorker if kind == "reference" else _PortWorker)(system, field, n, seed, n_warm)
        conn.send(("ready", w.S, len(w.rays)))
        while True:
            cmd = conn.recv()
            if cmd == "quit":
                break
            conn.send(("done", w.step(warm=(cmd == "warm"))))
    except Exception as e:            # noqa: BLE001 -- reported to the parent
        conn.send(("error", repr(e)))


def reference_available():
    import ref_shim
    return ref_shim.available()


def run(system="double_gauss", field=(0., .7), rays_per_proc=400000, procs=None, steps=1,
        warmup=1, kind=None, warm_rays=20000, rays_total=None):
    """-> dict(value ray-surfaces/s, kind, cores, seconds per timed step, ...)
    `rays_total`: shard a bundle of that many rays per wavelength over the
    processes (rays_per_proc = ceil(rays_total/procs)) -- the actual workload,
    ray-sharded over all cores, instead of a sample"""
    procs = procs or os.cpu_count() or 1
    if rays_total:
        rays_per_proc = -(-int(rays_total)//procs)
    if kind is None:
        kind = "reference" if reference_available() else "port"
    try:                                   # bound the resident set: ~1.3 kB per ray
        import psutil
        cap = int(psutil.virtual_memory().available*0.5/procs/1400)
        rays_per_proc = max(1000, min(rays_per_proc, cap))
    except ImportError:
        pass
    ctx = mp.get_context("fork")
    conns, ps = [], []
    for r in range(procs):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_worker, args=(b, kind, system, tuple(field), rays_per_proc,
                                              77 + r, warm_rays), daemon=True)
        p.start()
        conns.append(a)
        ps.append(p)

    def collect():
        out = []
        for c in conns:
            msg = c.recv()
            if msg[0] == "error":
                raise RuntimeError("cpu_bench worker failed: " + msg[1])
            out.append(msg)
        return out
    S, nl = collect()[0][1:3]

    def one(cmd):
        t0 = time.perf_counter()
        for c in conns:
            c.send(cmd)
        collect()
        return time.perf_counter() - t0
    # warm-up: the first step is full size (first touch of the result arrays,
    # allocator warm), further ones trace `warm_rays`-ray bundles
    for k in range(warmup):
        one("step" if k == 0 else "warm")
    secs = [one("step") for _ in range(steps)]
    for c in conns:
        c.send("quit")
    for p in ps:
        p.join(5)
    n = rays_per_proc*procs
    per_step = n*nl*S
    return {"value": per_step*steps/sum(secs), "kind": kind, "cores": procs,
            "rays_per_step_and_wavelength": n, "rays_per_proc": rays_per_proc,
            "wavelengths": nl, "surfaces": S, "steps": steps, "warmup": warmup,
            "seconds": secs, "ray_surfaces_per_step": per_step}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--system", default="double_gauss")
    ap.add_argument("--field", type=float, nargs=2, default=(0., .7))
    ap.add_argument("--rays-per-proc", type=int, default=400000)
    ap.add_argument("--rays-total", type=int, default=0)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--kind", default=None, choices=[None, "reference", "port"])
    a = ap.parse_args()
    print(json.dumps(run(a.system, a.field, a.rays_per_proc, a.procs or None, a.steps,
                         a.warmup, a.kind, rays_total=a.rays_total or None)))


if __name__ == "__main__":
    main()
