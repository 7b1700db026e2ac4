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
pg.close()
dist.barrier()
dist.destroy_process_group()
eng.close()
