#!/usr/bin/env python
"""Benchmark of the geometric propagate hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (N=1, and per rank for N>1 -- weak scaling): BASELINE.json configs[1],
"Double-Gauss 12-surface, 1e7 rays, 3 wavelengths, FP64, 1xB200": one STEP is
one pass of the hot path over the three wavelength bundles (3 launches of the
trace kernel, 1e7 aimed rays x 12 surfaces each, clip=True, full trace
y,u,i,t stored).  Metric: ray-surface intersections per second.

  value     device-resident inputs and outputs, CUDA events on the launching
            stream around exactly K steps (max over ranks)
  e2e       the same work through the host-buffer C-ABI call rtx_trace_host:
            H2D of the launch rays and D2H of the whole trace inside the
            timed region (pinned host buffers)
  roofline  HBM: algorithmic bytes N*(6w + 10w*S) per launch / mean launch
            duration (CUDA events around every launch, separate pass)
  cpu_baseline  the numpy oracle port of the reference path on the host cores,
            bounded sample

`--impl reference` times the reference's CPU path (the numpy port in oracle/,
the reference itself being pure Python that cannot travel to the GPU box) on
all host cores, on a bounded sample of the same workload per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ray-surface intersections/sec"
UNIT = "ray-surfaces/s"
SYSTEM = "double_gauss"
FIELD_INDEX = 3            # field (0, 0.7) in tests/golden/systems.json
N_RAYS = 10_000_000        # per wavelength
WORKLOAD = ("C2: Double-Gauss 12-surface, 1e7 rays x 3 wavelengths, FP64, "
            "clip=True, field (0,0.7), full trace (y,u,i,t) stored")


def load_system(name):
    from rayopt_b200.surface_table import table_from_json
    with open(os.path.join(ROOT, "tests", "golden", "systems.json")) as f:
        ent = json.load(f)[name]
    ent["tables"] = [table_from_json(t) for t in ent["tables"]]
    return ent


def make_rays(ent, li, n, seed):
    from rayopt_b200.rays import aim_infinite, disc
    aim = ent["aim"][li][FIELD_INDEX]
    return aim_infinite(aim["field"], disc(n, seed), aim["z"], aim["p"],
                        ent["object_angle"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_first(self, timeout=5.0):
        """block until nvidia-smi has produced its first sample"""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        # samples inside the timed region; a region shorter than the sampling
        # period falls back to the samples nearest to it (still under load:
        # warm-up before, roofline pass after)
        rows = [r for t, r in self.rows if t0 <= t <= t1]
        if len(rows) < 2:
            near = sorted(self.rows, key=lambda tr: min(abs(tr[0] - t0), abs(tr[0] - t1)))
            rows = [r for _, r in near[:3]]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
                power.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown",
                                "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_port(ent, n_per_proc, procs, repeat=1):
    """numpy oracle port of the reference path, ray-sharded over `procs`
    processes; returns (ray-surfaces/s, seconds, rays per wavelength)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_bench
    return cpu_bench.run(ent, FIELD_INDEX, n_per_proc, procs, repeat)


def run_reference(args):
    """--impl reference: the reference's CPU path (numpy port) on all cores"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ent = load_system(SYSTEM)
    cores = os.cpu_count() or 1
    n_per_proc = 50000
    for _ in range(args.warmup):
        cpu_port(ent, 2000, cores)
    tot_rs, dt = 0.0, 0.0
    for _ in range(args.steps):
        rate, secs, n = cpu_port(ent, n_per_proc, cores)   # secs: the trace only
        tot_rs += n*3*ent["S"]
        dt += secs
    value = tot_rs/dt
    sample = "%d rays x 3 wavelengths x %d surfaces per step (of 1e7), %d processes" % (
        n_per_proc*cores, ent["S"], cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt/args.steps*1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays", type=int, default=N_RAYS)
    ap.add_argument("--exact", type=int, default=0, help="1: RTX_EXACT arithmetic")
    ap.add_argument("--direct", type=int, default=0, help="1: per-thread stores")
    ap.add_argument("--rpt", type=int, default=0, help="rays per thread (0: library default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from rayopt_b200.engine import Engine
    eng = Engine(local)
    ent = load_system(SYSTEM)
    S, nl, N = ent["S"], len(ent["tables"]), args.rays
    ld = ((N + 63)//64)*64
    w = 8

    # ---- device-resident workload: 3 bundles, 3 full result sets ----------
    host_rays = []
    dev = []
    for li in range(nl):
        y0, u0 = make_rays(ent, li, N, seed=1000*rank + li)
        host_rays.append((y0, u0))
        d = {"y0": eng.to_device(y0), "u0": eng.to_device(u0),
             "Y": eng.empty((S, ld, 3)), "U": eng.empty((S, ld, 3)),
             "I": eng.empty((S, ld, 3)), "T": eng.empty((S, ld))}
        dev.append(d)

    def step():
        for li in range(nl):
            d = dev[li]
            eng.trace_device(ent["tables"][li], d["y0"], d["u0"], d["Y"], d["U"], d["I"],
                             d["T"], N=N, ld=ld, clip=True, exact=bool(args.exact),
                             direct=bool(args.direct), rpt=args.rpt)

    def barrier():
        eng.sync()
        if dist is not None:
            dist.barrier()
            import torch
            torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step()
    if sampler:
        sampler.wait_first()
        for _ in range(3):          # keep the GPU under load while the sampler spins up
            step()
    barrier()
    l0 = eng.launch_count()
    t_wall0 = time.time()
    eng.timer_start()
    for _ in range(args.steps):
        step()
    ms = eng.timer_stop()
    t_wall1 = time.time()
    launches = eng.launch_count() - l0
    for _ in range(3):              # samples right after the region are still under load
        step()
    barrier()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    if dist is not None:
        import torch
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms/args.steps
    value = world*nl*N*S/(ms_per_step*1e-3)

    # ---- roofline of the dominant (only) kernel: events around each launch --
    per_launch = []
    for _ in range(2):
        for li in range(nl):
            d = dev[li]
            eng.trace_device(ent["tables"][li], d["y0"], d["u0"], d["Y"], d["U"], d["I"],
                             d["T"], N=N, ld=ld, clip=True, exact=bool(args.exact),
                             direct=bool(args.direct), rpt=args.rpt)
            per_launch.append(eng.last_kernel_ms())
    k_ms = statistics.mean(per_launch)
    alg_bytes = N*(6*w + 10*w*S)
    # DRAM traffic per launch: dram__bytes_read.sum + dram__bytes_write.sum of
    # this kernel at this size from ncu (profiles/r1_v8_dram_bytes_full_size.csv:
    # 0.481 GB read + 9.536 GB written; the last ~64 MB of results are still
    # in L2 when the kernel ends).  Only valid for the default workload.
    traffic = 10_017_000_000 if (N == N_RAYS and not args.direct) else None
    achieved = alg_bytes/(k_ms*1e-3)/1e9
    peak, peak_src = peaks()

    # samples of the timed results, checked against the oracle in the
    # cpu_baseline leg below (the only place bench.py touches oracle/)
    idx = np.arange(0, N, max(1, N//2000))[:2000]
    got_dev = np.stack([dev[0]["Y"].rows(j).download()[0][idx] for j in range(S)])

    for d in dev:
        for a in d.values():
            a.free()

    # ---- e2e: the call a user makes -- GeometricTrace.propagate() ----------
    # host (pinned) result arrays in the reference layout; inside the timed
    # region: H2D of the launch rays, the kernel, D2H of the whole trace.
    # The system is unrotated, so the drop-in stores u and i as two views of
    # one buffer (i[j] == u[j-1] bit for bit) and moves 56 B per ray-surface;
    # "full_copy" is the same through rtx_trace_host with all four arrays
    # (80 B per ray-surface).
    e2e = None
    got_e2e = None
    if not args.no_e2e:
        from rayopt_b200 import GeometricTrace, PackedSystem
        ps = PackedSystem(ent["wavelengths"], ent["tables"], [n[0] for n in ent["n"]])
        # one trace object (7.5 GB of page-locked result arrays), propagated
        # once per wavelength and step: the launch rays of bundle 0, the
        # surface table of each wavelength
        g = GeometricTrace(ps, engine=eng, exact=bool(args.exact))
        g.rays_given(host_rays[0][0], host_rays[0][1], l=ent["wavelengths"][0])

        def e2e_step():
            for l in ent["wavelengths"]:
                g.l = l
                g.propagate(clip=True)

        def timed(fn, steps):
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
            barrier()
            dt = time.perf_counter() - t0
            if dist is not None:
                import torch
                t = torch.tensor([dt], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return dt
        e2e_steps = max(1, min(args.steps, 3))
        dt = timed(e2e_step, e2e_steps)
        # samples of what came back to the host (last wavelength traced)
        got_e2e = (g.y[1:, idx].copy(), g.i[1:, idx].copy(),
                   bool(np.array_equal(g.i[2:, idx], g.u[1:-1, idx], equal_nan=True)))
        e2e = {"value": world*nl*N*S*e2e_steps/dt, "unit": UNIT,
               "h2d_bytes_per_step": nl*N*6*w, "d2h_bytes_per_step": nl*N*S*7*w,
               "steps": e2e_steps, "ms_per_step": dt/e2e_steps*1e3,
               "api": "GeometricTrace.propagate(clip=True) -> rtx_trace_host: pinned host arrays, "
                      "chunked H2D/kernel/D2H pipeline; y,u,t copied back, i is a view of u"}
        del g
        import gc
        gc.collect()
    if not args.no_e2e and world == 1:
        # all four arrays through the C ABI
        out = {"y": eng.pinned_empty((S, N, 3)), "u": eng.pinned_empty((S, N, 3)),
               "i": eng.pinned_empty((S, N, 3)), "t": eng.pinned_empty((S, N))}
        pin = []
        for y0, u0 in host_rays:
            py, pu = eng.pinned_empty(y0.shape), eng.pinned_empty(u0.shape)
            py[:], pu[:] = y0, u0
            pin.append((py, pu))

        def full_step():
            for li in range(nl):
                eng.trace(ent["tables"][li], pin[li][0], pin[li][1], clip=True, out=out,
                          exact=bool(args.exact), rpt=args.rpt)
        fsteps = max(1, min(args.steps, 2))
        dt = timed(full_step, fsteps)
        e2e["full_copy"] = {"value": world*nl*N*S*fsteps/dt, "ms_per_step": dt/fsteps*1e3,
                            "d2h_bytes_per_step": nl*N*S*10*w,
                            "api": "rtx_trace_host with y,u,i,t host outputs"}

    # ---- e2e for a spot-diagram consumer (rayopt/analysis.py:269-280): the
    # trace stays in HBM (rayopt_b200.ResidentTrace semantics), per wavelength
    # the launch rays go up and only y[-1] comes back
    if not args.no_e2e and world == 1:
        from rayopt_b200._lib import check, ptr
        d_in = (eng.empty((N, 3)), eng.empty((N, 3)))
        d_full = {"Y": eng.empty((S, ld, 3)), "U": eng.empty((S, ld, 3)),
                  "I": eng.empty((S, ld, 3)), "T": eng.empty((S, ld))}
        h_spot = eng.pinned_empty((N, 3))

        def spot_step():
            for li in range(nl):
                check(eng.lib.rtx_memcpy_h2d(eng.ctx, d_in[0].ptr, ptr(pin[li][0]), pin[li][0].nbytes))
                check(eng.lib.rtx_memcpy_h2d(eng.ctx, d_in[1].ptr, ptr(pin[li][1]), pin[li][1].nbytes))
                eng.trace_device(ent["tables"][li], d_in[0], d_in[1], d_full["Y"], d_full["U"],
                                 d_full["I"], d_full["T"], N=N, ld=ld, clip=True,
                                 exact=bool(args.exact))
                check(eng.lib.rtx_memcpy_d2h(eng.ctx, ptr(h_spot), d_full["Y"].rows(S - 1).ptr,
                                             h_spot.nbytes))
            eng.sync()
        ssteps = max(1, min(args.steps, 5))
        dt = timed(spot_step, ssteps)
        e2e["spot_consumer"] = {"value": nl*N*S*ssteps/dt, "ms_per_step": dt/ssteps*1e3,
                                "h2d_bytes_per_step": nl*N*6*w, "d2h_bytes_per_step": nl*N*3*w,
                                "api": "device-resident full trace (ResidentTrace semantics): rays "
                                       "up, kernel, only y[-1] back"}
        for a in list(d_in) + list(d_full.values()):
            a.free()

    # ---- CPU baseline: numpy port of the reference path ------------------
    # (also the checker of the timed GPU results: same rays, same tables)
    cpu = None
    parity = None
    if rank == 0 and not args.no_cpu:
        cores = os.cpu_count() or 1
        n_per_proc = 50000
        rate, secs, n = cpu_port(ent, n_per_proc, cores, repeat=2)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d rays x 3 wavelengths x %d surfaces (of 1e7), %d processes, "
                         "%.1f s" % (n, S, cores, secs)}
        import np_oracle

        def rel(a, b):
            return float(np.nanmax(np.abs(a - b)/np.maximum(np.abs(b), 1.0)))
        want = np_oracle.trace(ent["tables"][0], host_rays[0][0][idx], host_rays[0][1][idx],
                               clip=True)
        parity = {"sample_rays": len(idx),
                  "device_resident": {
                      "nan_mask_equal": bool(np.array_equal(np.isnan(got_dev), np.isnan(want[0]))),
                      "max_rel_err_y": rel(got_dev, want[0])}}
        if got_e2e is not None:
            wl = np_oracle.trace(ent["tables"][nl - 1], host_rays[0][0][idx],
                                 host_rays[0][1][idx], clip=True)
            parity["e2e_host_arrays"] = {
                "nan_mask_equal": bool(np.array_equal(np.isnan(got_e2e[0]), np.isnan(wl[0])) and
                                       np.array_equal(np.isnan(got_e2e[1]), np.isnan(wl[2]))),
                "max_rel_err_y": rel(got_e2e[0], wl[0]), "max_rel_err_i": rel(got_e2e[1], wl[2]),
                "i_is_view_of_u": got_e2e[2]}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_wavelength": N, "surfaces": S,
                       "wavelengths": nl, "parallelism": "rays sharded x%d" % world,
                       "arithmetic": "exact" if args.exact else "fast",
                       "stores": "direct" if args.direct else "tma-bulk",
                       "kernel_config": "rpt=%s store=%s warps=%s nbuf=%s (0/unset: library default rpt 2, per-CTA TMA bulk stores, 16 warps, 1 staging buffer)" % (args.rpt, os.environ.get("RTX_STORE", "-"), os.environ.get("RTX_WARPS", "-"), os.environ.get("RTX_NBUF", "-")),
                       "l2": "outputs %.1f GB per launch >> 126 MB L2 (no flush needed)"
                             % (alg_bytes/1e9)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved/peak, "traffic": traffic,
                         "traffic_source": "ncu dram__bytes_read.sum+dram__bytes_write.sum per launch, "
                                           "profiles/r1_v8_dram_bytes_full_size.csv",
                         "peak_source": peak_src,
                         "kernel": "rtx::trace_kernel<double>", "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": alg_bytes},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "parity_check": parity,
        }))
    if dist is not None:
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
