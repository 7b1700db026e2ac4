#!/usr/bin/env python
"""Run under torchrun on N GPUs: ray-sharded trace + NCCL all-gather of the
last-surface intercepts and all-reduced rms, checked against the oracle.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tests/gpu_scripts/multi_gpu_check.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, torch.distributed as dist
import np_oracle, bench
from rayopt_b200.engine import Engine
from rayopt_b200.sharding import PeerGather, ShardedTrace, TorchComm

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Engine(local)
comm = TorchComm(dist)
ent = bench.load_system("double_gauss")
n = 2_000_003
y0, u0 = bench.make_rays(ent, 0, n, 0)
st = ShardedTrace(comm, engine=eng)
t0 = time.perf_counter()
spot = st.spot(ent["tables"][0], y0, u0, clip=True)
t1 = time.perf_counter()
rms = st.rms(ent["tables"][0], y0, u0, clip=False)
idx = np.arange(0, n, 997)
want = np_oracle.trace(ent["tables"][0], y0[idx], u0[idx], clip=True)[0][-1]
ok = np.array_equal(np.isnan(spot[idx]), np.isnan(want)) and np.nanmax(np.abs(spot[idx] - want)) < 1e-9
full = np_oracle.trace(ent["tables"][0], y0[::50], u0[::50], clip=False)[0][-1]
print("rank %d/%d: gathered spot %s in %.3f s, sample parity %s, rms %.12g" % (
    comm.rank, comm.world, spot.shape, t1 - t0, ok, rms), flush=True)
assert ok
# fused trace + gather over NVLink peer memory (one kernel per rank, no NCCL in the data path)
pg = PeerGather(eng, dist, n)
d_y0, d_u0 = eng.to_device(pg.local(y0)), eng.to_device(pg.local(u0))
pg.spot(ent["tables"][0], d_y0, d_u0, clip=True)          # warm-up (IPC mappings, peer access)
dist.barrier()
t0 = time.perf_counter()
spot2 = pg.spot(ent["tables"][0], d_y0, d_u0, clip=True)
t2 = time.perf_counter() - t0
kms = eng.last_kernel_ms()
same = np.array_equal(spot2, spot, equal_nan=True)
print("rank %d/%d: FUSED peer-store gather %s in %.4f s (kernel %.3f ms incl. NVLink stores), "
      "identical to the NCCL all-gather result: %s" % (comm.rank, comm.world, spot2.shape, t2, kms, same),
      flush=True)
assert same
# the statistics path: every rank reduces its shard inside the trace kernel
# (rtx_trace_reduce), ONE NCCL all-reduce of 20 doubles (Engine-side moments,
# TorchComm.sum) -> rms / refocus shift of the whole bundle
sub = np.isfinite(spot[:, 0])
y0s, u0s = y0[sub], u0[sub]                         # the reference's rms is not NaN-masked
chief = np_oracle.trace(ent["tables"][0], y0s[:1], u0s[:1], clip=True)
center = np.r_[chief[0][-1, 0, :2], chief[2][-1, 0, :2]/chief[2][-1, 0, 2]]
rms_f = st.rms_fused(ent["tables"][0], y0s, u0s, clip=True, center=center)
m = st.moments(ent["tables"][0], y0s, u0s, clip=True, center=center)
ys = spot[sub][:, :2]
want_rms = np.sqrt(np.square(ys - ys.mean(0)).sum(1).mean())
print("rank %d/%d: fused rms %.15g (gathered spot: %.15g), %d rays reduced over %d ranks" % (
    comm.rank, comm.world, rms_f, want_rms, int(m[5]), comm.world), flush=True)
assert m[5] == len(y0s) and abs(rms_f - want_rms) < 1e-11*want_rms
# gather of y[-1] AND i[-1] (through-focus spots, analysis.py:274-280)
bi = eng.empty((pg.npad, 3))
handles = [None]*comm.world
dist.all_gather_object(handles, eng.ipc_export(bi))
ptrs_i = [bi.ptr if r == comm.rank else eng.ipc_open(h) for r, h in enumerate(handles)]
n_local = pg.b[comm.rank + 1] - pg.b[comm.rank]
eng.trace_gather(ent["tables"][0], d_y0, d_u0, pg.ptrs, pg.b[comm.rank], N=n_local, clip=True,
                 dst_i_ptrs=ptrs_i)
eng.sync()
dist.barrier()
inc = bi.download()[:n]
want_i = np_oracle.trace(ent["tables"][0], y0[idx], u0[idx], clip=True)[2][-1]
ok_i = np.array_equal(np.isnan(inc[idx]), np.isnan(want_i)) and np.nanmax(np.abs(inc[idx] - want_i)) < 1e-12
print("rank %d/%d: gathered i[-1] parity %s" % (comm.rank, comm.world, ok_i), flush=True)
assert ok_i
dist.barrier()
for r, q in enumerate(ptrs_i):
    if r != comm.rank:
        eng.ipc_close(q)
pg.close()
dist.barrier()
dist.destroy_process_group()
eng.close()
