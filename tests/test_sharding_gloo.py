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
        def reducer(table, y, u, clip, w, center):
            """numpy stand-in for rtx_trace_reduce (same 20 sums)"""
            Y, U, I, T = np_oracle.trace(table, y, u, clip=clip)
            cy, cu = np.asarray(center[:2]), np.asarray(center[2:])
            ww = np.ones(len(y)) if w is None else w
            d = Y[-1][:, :2] - cy
            s = I[-1][:, :2]/I[-1][:, 2:] - cu
            f, g = np.isfinite(d).all(1), np.isfinite(s).all(1)
            return np.r_[ww[f].sum(), (ww[f, None]*d[f]).sum(0), (ww[f]*(d[f]**2).sum(1)).sum(),
                         f.sum(), len(y), d[f].sum(0),
                         g.sum(), d[g].sum(0), s[g].sum(0), ww[g].sum(), (ww[g, None]*d[g]).sum(0),
                         (ww[g, None]*s[g]).sum(0), (ww[g]*(d[g]*s[g]).sum(1)).sum(),
                         (ww[g]*(s[g]**2).sum(1)).sum()]
        st = ShardedTrace(TorchComm(dist), tracer=tracer, reducer=reducer)
        n = c["y0"].shape[0] - 5                     # uneven shards
        spot = st.spot(c["table"], c["y0"][:n], c["u0"][:n], clip=False)
        rms = st.rms(c["table"], c["y0"][:n], c["u0"][:n], clip=False)
        center = np.r_[c["Y"][-1, 0, :2], c["I"][-1, 0, :2]/c["I"][-1, 0, 2]]
        m = st.moments(c["table"], c["y0"][:n], c["u0"][:n], clip=False, center=center)
        rms2 = st.rms_fused(c["table"], c["y0"][:n], c["u0"][:n], clip=False, center=center)
        w = np.linspace(1, 2, n)/n
        rms3 = st.rms_fused(c["table"], c["y0"][:n], c["u0"][:n], w=w, clip=False, center=center)
        q.put((rank, spot, rms, m, rms2, rms3))
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
    from rayopt_b200.engine import Engine
    i = c["I"][-1, :n]
    u = i[:, :2]/i[:, 2:]
    yy, uu = want[:, :2] - want[:, :2].mean(0), u - u.mean(0)
    shift = -(yy*uu).sum()/(uu*uu).sum()
    w = np.linspace(1, 2, n)/n
    for rank, spot, rms, m, rms2, rms3 in res:
        assert spot.shape == (n, 3)
        assert np.array_equal(spot, want)
        assert abs(rms - np_oracle.rms(want)) < 1e-14
        # the all-reduced one-pass moments: counts, rms (default and given
        # weights) and the refocus shift of the whole bundle
        assert m[5] == n and m[4] == n and m[8] == n
        assert abs(rms2 - np_oracle.rms(want)) < 1e-13
        assert abs(rms3 - np_oracle.rms(want, w)) < 1e-13
        assert abs(Engine.focus_shift_from_moments(m) - shift) < 1e-10*abs(shift)


def test_bounds_cover_everything():
    from rayopt_b200.sharding import bounds
    for n in (0, 1, 7, 64, 1000003):
        for g in (1, 2, 3, 8):
            b = bounds(n, g)
            assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
            assert max(y - x for x, y in zip(b, b[1:])) - min(y - x for x, y in zip(b, b[1:])) <= 1
