"""torchrun -N: host <-> device copy bandwidth per rank with all ranks copying at once (explains the e2e numbers at N > 1)."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
MB = 1 << 20
hin = torch.empty(160 * MB, dtype=torch.uint8).pin_memory()
hout = torch.empty(320 * MB, dtype=torch.uint8).pin_memory()
din = torch.empty(160 * MB, dtype=torch.uint8, device=dev)
dout = torch.empty(320 * MB, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()

def run(name, h2d, d2h, reps=10):
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                din[: h2d * MB].copy_(hin[: h2d * MB], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                hout[: d2h * MB].copy_(dout[: d2h * MB], non_blocking=True)
    s1.synchronize(); s2.synchronize()
    dt = (time.perf_counter() - t0) / reps
    barrier()
    print(f"rank {rank}/{world} {name}: {dt * 1e3:.2f} ms  h2d {h2d / 1024 / dt:.1f} GB/s  d2h {d2h / 1024 / dt:.1f} GB/s", flush=True)

run("h2d 160MB alone", 160, 0)
run("d2h 320MB alone", 0, 320)
run("h2d 160 + d2h 320 together", 160, 320)
run("h2d 80 + d2h 160 together (the per-rank share at N=2)", 80, 160)
if world > 1:
    dist.destroy_process_group()
