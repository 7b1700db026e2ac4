"""CPU timing of the oracle port (TEST/BENCH INFRASTRUCTURE: used only by
bench.py's cpu_baseline and --impl reference legs).

The reference's geometric trace is single-threaded numpy; "all host cores" is
obtained the way SURVEY.md 6 did it: ray-sharded multiprocessing, every process
tracing its own bundle through all wavelengths with np_oracle.trace.
"""
import multiprocessing as mp
import os
import sys
import time


HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _work(args):
    tables, rays, repeat = args
    import np_oracle
    t0 = time.perf_counter()
    for _ in range(repeat):
        for table, (y, u) in zip(tables, rays):
            np_oracle.trace(table, y, u, clip=True)
    return time.perf_counter() - t0


def run(ent, field_index, n_per_proc, procs, repeat=1):
    from rayopt_b200.rays import aim_infinite, disc
    jobs = []
    for p in range(procs):
        rays = []
        for li in range(len(ent["tables"])):
            aim = ent["aim"][li][field_index]
            rays.append(aim_infinite(aim["field"], disc(n_per_proc, 77 + p), aim["z"],
                                     aim["p"], ent["object_angle"]))
        jobs.append((ent["tables"], rays, repeat))
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        pool.map(_work, [(ent["tables"], jobs[0][1][:1], 1)]*procs)   # warm the workers
        t0 = time.perf_counter()
        pool.map(_work, jobs)
        dt = time.perf_counter() - t0
    n = n_per_proc*procs
    rs = n*len(ent["tables"])*ent["S"]*repeat
    return rs/dt, dt, n
