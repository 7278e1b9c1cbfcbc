"""Probe what peer-memory plumbing the box offers (run under torchrun, N >= 2)."""
import os, sys, time
import torch, torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if rank == 0:
    print("p2p", [[torch.cuda.can_device_access_peer(i, j) if i != j else None for j in range(world)] for i in range(world)], flush=True)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    hdl = symm.rendezvous(t, dist.group.WORLD)
    print(rank, "symm ok", "bufs", [hex(p) for p in hdl.buffer_ptrs], "sig", [hex(p) for p in hdl.signal_pad_ptrs],
          "mc", hex(hdl.multicast_ptr) if hdl.multicast_ptr else None, "sigsize", hdl.signal_pad_size, flush=True)
except Exception as ex:
    print(rank, "symm failed", repr(ex)[:500], flush=True)
dist.barrier()
dist.destroy_process_group()
