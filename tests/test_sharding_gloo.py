"""World-size-2 gloo test of the N>1 host path (CPU): shard boundaries, the
all-gather of last-surface intercepts and the all-reduced rms.  The tracer is
the oracle here (no GPU in this container); on the GPU box the same class runs
on the CUDA engine (tests/test_gpu_parity.py::test_sharded_trace_single_rank)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import np_oracle
    from rayopt_b200.sharding import ShardedTrace, TorchComm
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = load_golden("double_gauss_f1_noclip")

        def tracer(table, y, u, clip, keep_last):
            Y, U, I, T = np_oracle.trace(table, y, u, clip=clip)
            return (Y[-1:], U[-1:], I[-1:], T[-1:]) if keep_last else (Y, U, I, T)
        st = ShardedTrace(TorchComm(dist), tracer=tracer)
        n = c["y0"].shape[0] - 5                     # uneven shards
        spot = st.spot(c["table"], c["y0"][:n], c["u0"][:n], clip=False)
        rms = st.rms(c["table"], c["y0"][:n], c["u0"][:n], clip=False)
        q.put((rank, spot, rms))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gather_and_rms():
    import np_oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    c = load_golden("double_gauss_f1_noclip")
    n = c["y0"].shape[0] - 5
    want = c["Y"][-1, :n]
    for rank, spot, rms in res:
        assert spot.shape == (n, 3)
        assert np.array_equal(spot, want)
        assert abs(rms - np_oracle.rms(want)) < 1e-14


def test_bounds_cover_everything():
    from rayopt_b200.sharding import bounds
    for n in (0, 1, 7, 64, 1000003):
        for g in (1, 2, 3, 8):
            b = bounds(n, g)
            assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
            assert max(y - x for x, y in zip(b, b[1:])) - min(y - x for x, y in zip(b, b[1:])) <= 1
