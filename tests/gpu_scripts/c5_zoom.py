#!/usr/bin/env python
"""BASELINE config C5: 20-surface zoom lens, 5 fields x 5 wavelengths x 1e7
rays, FP64, the 25 bundles distributed over the GPUs of one box (whole
(field, wavelength) bundles per GPU).  Under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29571 tests/gpu_scripts/c5_zoom.py

Each bundle: launch rays generated in HBM (hexapolar, the reference's aim
solution for that field and wavelength), full trace (y,u,i,t of all 20
surfaces stored, 16.5 GB), CUDA-event kernel time; a sample of the last bundle
of every rank is checked against the oracle."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, torch.distributed as dist
import np_oracle, bench
from rayopt_b200.engine import Engine

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
eng = Engine(local)
ent = bench.load_system("zoom")
S = ent["S"]
NR = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
field_idx = [0, 1, 2, 4, 5]                          # fields 0, .25, .5, .75, 1
bundles = [(fi, li) for fi in field_idx for li in range(5)]
mine = bundles[rank::world]
rings = int(np.sqrt(NR/3. - 1/12.) - 1/2.)
N = 1 + 3*rings*(rings + 1)
ld = (N + 63)//64*64
Y, U, I = (eng.empty((S, ld, 3)) for _ in range(3))
T = eng.empty((S, ld))
total_ms, last = 0.0, None
for rep in range(2):                                 # first pass warms up
    total_ms = 0.0
    for fi, li in mine:
        aim = ent["aim"][li][fi]
        y0, u0 = eng.aim_infinite_device(aim["field"], aim["z"], aim["p"], ent["object_angle"],
                                         nrays=NR)
        eng.trace_device(ent["tables"][li], y0, u0, Y, U, I, T, N=N, ld=ld, clip=True)
        total_ms += eng.last_kernel_ms()
        if last is not None:
            last[0].free(), last[1].free()
        last = (y0, u0, fi, li)
    dist.barrier()
t = torch.tensor([total_ms], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
worst = float(t.item())
y0, u0, fi, li = last
idx = np.arange(0, N, N//1500)
hy, hu = y0.download()[idx], u0.download()[idx]
want = np_oracle.trace(ent["tables"][li], hy, hu, clip=True)
got = np.stack([Y.rows(j).download()[0][idx] for j in range(S)])
ok = bool(np.array_equal(np.isnan(got), np.isnan(want[0])) and
          np.nanmax(np.abs(got - want[0])/np.maximum(np.abs(want[0]), 1.0)) < 1e-10)
print(json.dumps({"rank": rank, "world": world, "bundles_total": len(bundles), "bundles_this_rank": len(mine),
                  "rays_per_bundle": N, "surfaces": S, "kernel_ms_this_rank": total_ms,
                  "kernel_ms_max_over_ranks": worst,
                  "ray_surfaces_per_s_total": len(bundles)*N*S/(worst*1e-3),
                  "per_gpu_GBps": len(mine)*N*(48 + 80*S)/(total_ms*1e-3)/1e9, "parity_ok": ok}), flush=True)
assert ok
dist.barrier()
dist.destroy_process_group()
eng.close()
