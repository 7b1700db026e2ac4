"""Ray sharding over the GPUs of one box.

The path shards naturally: rays are independent, the surface table is
replicated, there is NO exchange inside the surface loop.  One process per
GPU; rank r owns the contiguous ray range ``bounds(N, G)[r]``.  The only
collectives the path has are optional epilogues:

* ``gather_last``  -- all-gather of last-surface intercepts ``y[-1]`` when a
  caller needs the full spot on every rank (24 B/ray FP64);
* ``rms``          -- all-reduce of 8 FP64 moments (rtx_moments) instead,
  when only statistics are needed.

``torch.distributed`` is plumbing only (rendezvous, NCCL/gloo collectives on
buffers the engine owns); the engine itself is torch-free.  The tracer is
injected so that the host logic is testable on CPU with gloo.
"""
import numpy as np


def bounds(n, world, align=1):
    """contiguous shard boundaries: rank r owns [b[r], b[r+1]); interior
    boundaries are multiples of `align` rays"""
    b = [((r*n)//world)//align*align for r in range(world)] + [n]
    return b


def shard(a, rank, world, align=1):
    b = bounds(len(a), world, align)
    return a[b[rank]:b[rank + 1]]


class TorchComm:
    """all-gather / all-reduce of numpy arrays over torch.distributed
    (backend nccl on GPUs, gloo on CPU)."""

    def __init__(self, dist=None, device=None):
        if dist is None:
            import torch.distributed as dist
        import torch
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if device is None:
            device = "cuda" if dist.get_backend() == "nccl" else "cpu"
        self.device = device

    def sum(self, v):
        t = self.torch.as_tensor(np.asarray(v, np.float64), device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def max(self, v):
        t = self.torch.as_tensor(np.asarray(v, np.float64), device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.cpu().numpy()

    def all_gather_rows(self, local, n_total):
        """concatenate per-rank (n_r, k) blocks (contiguous `bounds` shards)
        into (n_total, k) on every rank; shards are padded to equal length for
        the collective (all_gather_into_tensor needs equal sizes)."""
        torch, dist = self.torch, self.dist
        b = bounds(n_total, self.world)
        per = max(b[r + 1] - b[r] for r in range(self.world))
        k = local.shape[1]
        send = torch.zeros((per, k), dtype=torch.float64, device=self.device)
        send[:local.shape[0]] = torch.as_tensor(np.ascontiguousarray(local), device=self.device)
        recv = torch.empty((self.world*per, k), dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(recv, send)
        recv = recv.cpu().numpy().reshape(self.world, per, k)
        return np.concatenate([recv[r, :b[r + 1] - b[r]] for r in range(self.world)])

    def barrier(self):
        self.dist.barrier()


class LocalComm:
    """world of one"""
    rank, world = 0, 1

    def sum(self, v):
        return np.asarray(v, np.float64)

    max = sum

    def all_gather_rows(self, local, n_total):
        return np.asarray(local)

    def barrier(self):
        pass


def rms_from_moments(m, unit_weights=False):
    """see Engine.rms_from_moments (pure numpy; kept importable without CUDA)"""
    from .engine import Engine
    return Engine.rms_from_moments(m, unit_weights)


class ShardedTrace:
    """Trace a ray bundle sharded over ranks.

    `tracer(table, y0, u0, clip, keep_last)` -> (Y, U, I, T) host arrays for
    the LOCAL shard; by default the CUDA engine of this rank's GPU.
    `reducer(table, y0, u0, clip, w, center)` -> the 20 moments of
    rtx_trace_reduce for the LOCAL shard (host arrays in); by default the fused
    CUDA epilogue (the shard is uploaded, no intercept is stored or read back).
    """

    def __init__(self, comm=None, tracer=None, engine=None, reducer=None):
        self.comm = comm or LocalComm()
        if (tracer is None or reducer is None) and engine is None:
            from .engine import default_engine
            engine = default_engine()
        if tracer is None:
            tracer = lambda t, y, u, clip, keep_last: engine.trace(  # noqa: E731
                t, y, u, clip=clip, keep_last=keep_last)
        if reducer is None:
            def reducer(table, y, u, clip, w, center):
                dy, du = engine.to_device(y), engine.to_device(u)
                dw = None if w is None else engine.to_device(w)
                try:
                    return engine.trace_reduce(table, dy, du, clip=clip, w=dw, center=center)
                finally:
                    for a in (dy, du, dw):
                        if a is not None:
                            a.free()
        self.tracer = tracer
        self.reducer = reducer

    def moments(self, table, y0, u0, w=None, clip=False, center=None):
        """the 20 rms / refocus moments (include/rtx.h rtx_trace_reduce) of the
        WHOLE bundle: every rank reduces its shard inside the trace kernel, ONE
        all-reduce of 20 doubles adds them up.  `center` must be
        the same on every rank (e.g. the chief ray's intercept and slope)."""
        wl = None if w is None else self.local(np.asarray(w, float))
        m = self.reducer(table, self.local(y0), self.local(u0), clip, wl, center)
        return self.comm.sum(m)

    def rms_fused(self, table, y0, u0, w=None, clip=False, center=None):
        """GeometricTrace.rms of the sharded bundle from one fused launch per
        rank and one 160-byte all-reduce"""
        m = self.moments(table, y0, u0, w, clip, center)
        return rms_from_moments(m, unit_weights=w is None)

    def local(self, a):
        return shard(a, self.comm.rank, self.comm.world)

    def spot(self, table, y0, u0, clip=False):
        """full (N, 3) last-surface intercepts on every rank: each rank traces
        its shard (keep-LAST), then one all-gather."""
        n = len(y0)
        Y, U, I, T = self.tracer(table, self.local(y0), self.local(u0), clip, True)
        return self.comm.all_gather_rows(Y[0], n)

    def rms(self, table, y0, u0, w=None, clip=False):
        """GeometricTrace.rms of the whole bundle from per-rank moments
        (two all-reduces of 8 doubles; no ray data crosses the link)."""
        n = len(y0)
        Y, U, I, T = self.tracer(table, self.local(y0), self.local(u0), clip, True)
        y = Y[0][:, :2]
        wl = np.full(len(y), 1.0/n) if w is None else self.local(np.asarray(w, float))
        fin = np.isfinite(y).all(1)
        m = self.comm.sum([fin.sum(), len(y), y[fin, 0].sum(), y[fin, 1].sum()])
        if m[0] != m[1]:
            return float("nan")            # the reference's rms is not NaN-masked
        c = m[2:4]/m[1]
        r = (wl*np.square(y - c).sum(1)).sum()
        return float(np.sqrt(self.comm.sum([r])[0]))


class PeerGather:
    """Fused trace + all-gather over NVLink peer memory (rtx_trace_gather).

    Every rank owns a (Npad, 3) gather buffer on its GPU, exports it through
    CUDA IPC, and opens the buffers of all peers.  ``spot()`` then runs ONE
    kernel per rank whose last-surface TMA bulk stores go straight into all
    `world` buffers -- no separate collective, the transfer overlaps the
    march tile by tile.  torch.distributed only carries the 64-byte handles
    and the final barrier."""

    ALIGN = 64

    def __init__(self, engine, dist, n_total, dtype=np.float64, xy=False):
        """`xy`: gather x,y only -- (n_total, 2) buffers, 16 instead of 24 bytes
        per ray over NVLink (what a spot diagram reads)"""
        self.eng, self.dist, self.xy = engine, dist, bool(xy)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world > 8:
            raise ValueError("at most 8 peers (one NVSwitch box)")
        self.n = int(n_total)
        self.npad = (self.n + self.ALIGN - 1)//self.ALIGN*self.ALIGN + self.ALIGN
        self.buf = engine.empty((self.npad, 2 if xy else 3), dtype)
        handles = [None]*self.world
        dist.all_gather_object(handles, engine.ipc_export(self.buf))
        self._opened = []
        self.ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.buf.ptr)
            else:
                p = engine.ipc_open(h)
                self._opened.append(p)
                self.ptrs.append(p)
        self.b = bounds(self.n, self.world, self.ALIGN)

    def local(self, a):
        return a[self.b[self.rank]:self.b[self.rank + 1]]

    def spot(self, table, y0_local_dev, u0_local_dev, clip=False, exact=False):
        """full last-surface intercepts (n_total, 3) in this rank's device
        buffer; returns it as a host array"""
        n_local = self.b[self.rank + 1] - self.b[self.rank]
        if n_local:
            self.eng.trace_gather(table, y0_local_dev, u0_local_dev, self.ptrs,
                                  self.b[self.rank], N=n_local, clip=clip, exact=exact,
                                  xy=self.xy)
        self.eng.sync()
        self.dist.barrier()          # every rank's stores have landed everywhere
        return self.buf.download()[:self.n]

    def close(self):
        self.dist.barrier()
        for p in self._opened:
            self.eng.ipc_close(p)
        self._opened = []
        self.buf.free()
