"""Host <-> device copy rates of GPU 0 with the pinned buffers allocated from each NUMA node in turn (the process is bound
to the node's CPUs before cudaHostAlloc, so first touch places the pages there). Explains the box-to-box spread of e2e."""
import glob, os, subprocess, time
import torch
print(subprocess.run("nvidia-smi topo -m | head -12; lscpu | grep -i -E 'numa|socket|^CPU\\(s\\)'", shell=True, capture_output=True, text=True).stdout, flush=True)
dev = torch.device("cuda", 0)
torch.cuda.init()
bus = torch.cuda.get_device_properties(0)
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    busid = pynvml.nvmlDeviceGetPciInfo(h).busId
    busid = busid.decode() if isinstance(busid, bytes) else busid
    path = "/sys/bus/pci/devices/" + busid[-12:].lower() + "/numa_node"
    print("GPU 0", busid, "numa_node", open(path).read().strip() if os.path.exists(path) else "?", flush=True)
except Exception as ex:
    print("nvml:", ex)
MB = 1 << 20
din = torch.empty(160 * MB, dtype=torch.uint8, device=dev)
dout = torch.empty(320 * MB, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
all_cpus = os.sched_getaffinity(0)
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
def cpus_of(node):
    out = set()
    for part in open(node + "/cpulist").read().strip().split(","):
        if "-" in part:
            a, b = part.split("-"); out |= set(range(int(a), int(b) + 1))
        elif part:
            out.add(int(part))
    return out & all_cpus
cases = [("default", all_cpus)] + [(os.path.basename(n), cpus_of(n)) for n in nodes]
for name, cpus in cases:
    if not cpus:
        continue
    os.sched_setaffinity(0, cpus)
    hin = torch.empty(160 * MB, dtype=torch.uint8).pin_memory(); hin.fill_(1)
    hout = torch.empty(320 * MB, dtype=torch.uint8).pin_memory(); hout.fill_(1)
    def run(h2d, d2h, reps=10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
        s1.synchronize(); s2.synchronize()
        return (time.perf_counter() - t0) / reps
    a, b, c = run(1, 0), run(0, 1), run(1, 1)
    print(f"{name:8s} cpus={len(cpus):3d}: h2d 160MB {a * 1e3:.2f} ms ({0.16 / a:.1f} GB/s)  d2h 320MB {b * 1e3:.2f} ms ({0.3355 / b:.1f} GB/s)  both {c * 1e3:.2f} ms", flush=True)
    del hin, hout
os.sched_setaffinity(0, all_cpus)
