#!/usr/bin/env python
"""Is NVLink SHARP multicast (NVLS) reachable from this container?  Under
torchrun: allocates a symmetric-memory buffer, rendezvous, prints the
multicast pointer (0 = unsupported) and the peer pointers.  Feasibility probe
for a multicast variant of rtx_trace_gather (one store replicated by the
switch instead of one store per peer)."""
import os
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty((1 << 20,), dtype=torch.float64, device="cuda")
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
    print("rank %d: multicast_ptr=%#x buffer_ptrs=%s world=%d" % (
        dist.get_rank(), int(getattr(hdl, "multicast_ptr", 0) or 0),
        [hex(p) for p in hdl.buffer_ptrs], hdl.world_size), flush=True)
except Exception as e:                                      # noqa: BLE001
    print("rank %d: symmetric memory unavailable: %r" % (dist.get_rank(), e), flush=True)
dist.barrier()
dist.destroy_process_group()
