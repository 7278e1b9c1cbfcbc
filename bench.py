#!/usr/bin/env python
"""bench.py — the driver-facing benchmark (one JSON line on stdout from rank 0).

Workload (BASELINE.json configs[1]): 10 M random BSphere{Float32} leaves -> BBox{Float32} nodes,
UInt32 Morton, Int32 index: `BVH(leaves, BBox{Float32}; cache)` + `traverse(bvh; cache)` (LVT contact
detection). A "step" is one build + one contact traversal on the same 10 M synthetic spheres
(SURVEY.md §8d law, seed 42); metric = leaves/s = N / step time.

  value  : inputs resident in HBM, CUDA-event time on the launching stream, max over ranks.
  e2e    : same call sequence through the public API with HOST (pinned) input volumes and the contact
           list read back to pinned host memory inside the timed region.
  roofline: dominant kernel (the LVT traversal kernel), algorithmic bytes / CUDA-event duration measured
           live with the library's per-kernel event profiler.
  cpu_baseline / --impl reference: the reference ALGORITHM restated in C++ (oracle/, two-pass LVT,
           struct-moving stable sort, static thread partition) on the box's host cores. The real
           reference is a Julia package and cannot run in this image (no julia binary).

N > 1 (torchrun, one rank per GPU): the build does not shard — rank 0 builds and the tree is broadcast
(NCCL); the contact traversal shards by contiguous query-leaf ranges; shards are all-gathered so every
rank ends with the full contact list. Total work is fixed => "scaling": "strong".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_LEAVES = int(os.environ.get("IBVH_BENCH_N", 10_000_000))
SEED = 42
METRIC = "build+contact leaves/s @10M spheres"
UNIT = "leaves/s"
WORKLOAD = ("configs[1]: %d BSphere{Float32} leaves (centres U[0,1)^3, r = s(0.5+0.5u), s = 0.8124 N^(-1/3), seed 42), "
            "BBox{Float32} nodes, UInt32 Morton, Int32 index: BVH build (built_level=1) + LVT contact traversal (start_level=1)")


def bench_config():
    """The `config` object: identical in the GPU arm and in the reference arm (the driver compares them)."""
    return {"workload": WORKLOAD % N_LEAVES, "leaves": N_LEAVES}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm restated (oracle/) on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_step(O, np, leaves_template, nodes, contacts, counts, threads):
    """One build + two-pass LVT traversal on the CPU. Returns (seconds, num_contacts)."""
    import ctypes
    lv = leaves_template.copy()
    t0 = time.perf_counter()
    kind, fb, ibt, mb = O._desc(lv)
    f = np.float32
    mn, mx = np.zeros(3, f), np.zeros(3, f)
    rc = O.lib().orc_build(O._p(lv), len(lv), kind, fb, ibt, mb, O._p(nodes), O.BBOX, 4, 1, 1, O._p(mn), O._p(mx), threads, 100, 100, 100)
    assert rc == 0
    total = O.lib().orc_traverse_single(O._p(lv), len(lv), kind, fb, ibt, mb, O._p(nodes), O.BBOX, 4, 1, 1, O._p(contacts), len(contacts),
                                        O._p(counts), threads, 100)
    dt = time.perf_counter() - t0
    assert 0 <= total <= len(contacts), total
    return dt, int(total)


def cpu_setup(n):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle as O
    from ibvh_b200 import synth
    s = synth.random_spheres_np(n, seed=SEED)
    leaves = O.wrap(s)
    nodes = np.zeros(O.num_nodes(n), O.volume_dtype(O.BBOX, 4))
    contacts = np.zeros(int(6.0 * n) + 1024, O.pair_dtype(4))
    counts = np.zeros(n, np.int32)
    return O, np, leaves, nodes, contacts, counts


def cpu_calibrated_size(threads, budget_s):
    """Pick the sample size: run 200 k leaves, extrapolate (build+traverse is ~linear), cap at the full N."""
    O, np, leaves, nodes, contacts, counts = cpu_setup(200_000)
    cpu_step(O, np, leaves, nodes, contacts, counts, threads)
    dt, _ = cpu_step(O, np, leaves, nodes, contacts, counts, threads)
    rate = 200_000 / dt
    n = int(min(N_LEAVES, max(200_000, rate * budget_s * 0.8)))
    return n


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    budget = 150.0 / max(1, args.steps + args.warmup)
    n = cpu_calibrated_size(threads, min(budget, 20.0))
    O, np, leaves, nodes, contacts, counts = cpu_setup(n)
    for _ in range(args.warmup):
        cpu_step(O, np, leaves, nodes, contacts, counts, threads)
    t = 0.0
    total = 0
    for _ in range(args.steps):
        dt, total = cpu_step(O, np, leaves, nodes, contacts, counts, threads)
        t += dt
    ms = t / args.steps * 1e3
    value = n / (ms * 1e-3)
    sample = (f"{n}-leaf instance of the same law (full workload is {N_LEAVES}); C++17 restatement of the reference algorithm "
              f"(oracle/), {threads} threads, build + two-pass LVT; the Julia reference itself cannot run here")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(),
        "details": {"sample_leaves": n, "contacts_per_step": total},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def profile_rows(ib, handle):
    lib = ib.capi.lib()
    name = C.create_string_buffer(64)
    ms = C.c_float()
    rows = []
    for i in range(lib.ibvh_profile_count(handle)):
        lib.ibvh_profile_get(handle, i, name, 64, C.byref(ms))
        rows.append((name.value.decode(), float(ms.value)))
    return rows


def pair_checksum(torch, tensor_u8, count, itemsize=8):
    """Order-independent checksum of an IndexPair list on the device: (count, sum of keys, sum of squared keys), both
    sums wrapping mod 2^64. key = the pair's 8 bytes read as one int64 (Int32 pairs) or a * 2^32 + b (Int64 pairs)."""
    if count == 0:
        return (0, 0, 0)
    if itemsize == 8:
        k = tensor_u8[: count * 8].view(torch.int64)
    else:
        ab = tensor_u8[: count * 16].view(torch.int64).reshape(-1, 2)
        k = ab[:, 0] * (1 << 32) + ab[:, 1]
    return (int(count), int(k.sum().item()), int((k * k).sum().item()))


def timed_steps(torch, fn, steps, warm):
    """`warm` untimed + `steps` timed calls of fn(); CUDA events on the current stream. Returns (ms per step, last result)."""
    out = None
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ordered", action="store_true",
                    help="emit contacts in the reference's order (count -> scan -> write) instead of the one-pass unordered emission "
                         "(same set; the parity bar is the SORTED contact list)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rays", action="store_true", help="skip the secondary rays/s measurement")
    ap.add_argument("--workloads", default="ordered,bfs,pair,rebuild64,reference_shaped",
                    help="comma list of the extra BASELINE.json workloads to time after the headline (reported under `workloads`): "
                         "ordered (the headline step in the reference's contact order), pair (configs[2]: 5 M + 5 M BVH-vs-BVH, "
                         "partial and full build), rebuild64 (configs[4]: 100 M leaves, UInt64 / Int64, cached rebuild + contacts), "
                         "reference_shaped (the naive GPU proxy of the reference's CUDA.jl backend on the headline step), bfs (the headline step "
                         "with traverse(bvh, BFSTraversal()), single GPU). 'none' skips them")
    ap.add_argument("--no-defer", action="store_true", help="single GPU: synchronous traversal calls (the host reads each step's contact count before it enqueues the next build)")
    ap.add_argument("--gather", default="fused", choices=["fused", "peer", "nccl"],
                    help="N > 1: how the contact shards reach every rank. fused: the traversal kernel itself writes each contact "
                         "into every rank's list (slots from one counter on rank 0, multimem.st over NVLink; unordered only); "
                         "peer: traverse locally, then the library's all-gather kernel (ibvh_allgather_pairs); nccl: NCCL collectives")
    ap.add_argument("--build-mode", default="replicate", choices=["replicate", "broadcast"],
                    help="N > 1: every rank builds the (deterministic, bit-identical) tree itself, or rank 0 builds and the "
                         "tree is broadcast with NCCL. The build does not shard either way; replicate is faster as long as a "
                         "build (1.0 ms at 10 M leaves) costs less than moving 480 MB of leaves + nodes")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import ibvh_b200 as ib
    from ibvh_b200 import dist as ibdist
    from ibvh_b200 import synth

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    # stdout carries exactly ONE line, the JSON: everything libraries write to file descriptor 1 meanwhile (the NCCL version
    # banner, NCCL_DEBUG output the driver may have asked for) is sent to stderr; the JSON goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)      # (NCCL_DEBUG is left as the caller set it; its output lands on stderr)
    warm = max(3, args.warmup)
    n = N_LEAVES
    ordered = args.ordered
    extra = [] if args.workloads in ("", "none") else [w.strip() for w in args.workloads.split(",") if w.strip()]

    # ---- inputs resident in HBM (generated on device, identical on every rank) -----------------
    vols = synth.random_spheres_torch(n, dev, seed=SEED)
    src = ib.DeviceArray(vols.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
    bounds = ibdist.shard_bounds(n, world)
    qb, qe = bounds[rank]

    state = {"bvh": None, "tr": None, "full": None, "peer": None}
    fused = world > 1 and args.gather == "fused" and not ordered

    def step_device():
        """One step with inputs in HBM. Returns the number of contacts this rank holds at the end."""
        if world == 1:
            # the traversal is enqueued without a host round trip (defer) and the NEXT build is enqueued before the
            # host waits for this step's contact count: the GPU never idles between steps. Every step's count is
            # still read inside the timed region (one step late; the last one in finish_pending()).
            bvh = ib.BVH(src, ib.BBox(), cache=state["bvh"])
            prev = state["tr"]
            n_prev = prev.num_contacts if prev is not None else 0          # resolves the previous step's traversal
            tr = ib.traverse(bvh, cache=prev, ordered=ordered, defer=not ordered and not args.no_defer)
            state["bvh"], state["tr"] = bvh, tr
            return n_prev
        # the build does not shard: replicate it, or build on rank 0 and broadcast the tree (leaves + nodes);
        # then shard the traversal by query range and gather the shards
        if args.build_mode == "replicate":
            bvh = ib.BVH(src, ib.BBox(), cache=state["bvh"])
            state["bvh"] = bvh
        else:
            if rank == 0 or state["bvh"] is None:
                bvh = ib.BVH(src, ib.BBox(), cache=state["bvh"])        # every rank builds once to own the buffers
                state["bvh"] = bvh
            bvh = state["bvh"]
            dist.broadcast(bvh.leaves.tensor, src=0)
            dist.broadcast(bvh.nodes.tensor, src=0)
        if state["peer"] is not None and fused:
            tr = ib.traverse(bvh, cache=state["tr"], ordered=False, query_range=(qb, qe - qb), peer=state["peer"])
            state["full"] = tr.cache1.tensor
            return tr.num_contacts
        tr = ib.traverse(bvh, cache=state["tr"], ordered=ordered, query_range=(qb, qe - qb))
        state["tr"] = tr
        if state["peer"] is not None:
            full, total, _ = state["peer"].gather(tr.cache1.tensor, tr.num_contacts)
        else:
            full, counts = ibdist.gather_shards(tr.cache1.tensor, tr.num_contacts, tr.cache1.dtype.itemsize)
            total = int(sum(counts))
        state["full"] = full
        return total

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # first call sizes the caches; give cache1 head-room so later steps never regrow
    ncontacts = step_device()
    if world > 1 and args.gather in ("peer", "fused"):
        try:
            state["peer"] = ibdist.PeerGather(int(ncontacts * 1.10) + 8192 * world, 8, dev)
        except Exception as ex:                          # no symmetric memory on this box: NCCL collectives instead
            if rank == 0:
                print(f"[bench] peer memory unavailable ({ex!r}); gathering with NCCL", file=sys.stderr, flush=True)
            state["peer"], fused, args.gather = None, False, "nccl"
        if fused and not state["peer"].peer.multicast:
            fused = False                                # no NVLS multicast on this box: two-step peer gather
    if world == 1:
        tr = state["tr"]
        state["tr"] = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(int(tr.num_contacts * 1.05) + 1024, ib.pair_dtype(), dev), tr.cache2)
    for _ in range(warm):
        ncontacts = step_device()

    # ---- N > 1: the gathered contact list of the timed configuration against the single-GPU list, on every rank ----
    verify = None
    if world > 1:
        mc = bool(state["peer"] is not None and state["peer"].peer.multicast)
        print(f"[bench] rank {rank}/{world} device cuda:{local} ({torch.cuda.get_device_name(local)}) gather={args.gather} "
              f"fused={fused} multicast={'yes' if mc else 'no'} build={args.build_mode} shard=[{qb},{qe})", file=sys.stderr, flush=True)
        got = pair_checksum(torch, state["full"], int(ncontacts))
        single = ib.traverse(state["bvh"], ordered=False)                 # the whole problem on this rank alone
        want = pair_checksum(torch, single.cache1.tensor, single.num_contacts)
        ok = torch.tensor([1 if got == want else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        verify = {"gathered_equals_single_gpu_list": bool(ok.item()), "contacts": want[0], "checksum": [want[1], want[2]],
                  "what": "count + order-independent checksum (sum and sum of squares of the 64-bit pair keys, mod 2^64) of the gathered "
                          "list on EVERY rank against the same rank's own unsharded traversal, before the timed region"}
        if not ok.item():
            raise SystemExit(f"[bench] rank {rank}: gathered contact list {got} != single-GPU list {want}")
        del single

    def finish_pending():
        """Single GPU: the count of the last (deferred) traversal — read before the closing event is recorded."""
        return state["tr"].num_contacts if world == 1 else None

    finish_pending()
    sampler = ClockSampler(local)
    sync_all()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        ncontacts = step_device()
    if world == 1:
        ncontacts = finish_pending()
    e1.record()
    sync_all()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = n / (ms_step * 1e-3)

    # ---- per-kernel durations (CUDA events on the launching stream, inside the library) ---------
    handle = state["bvh"]._handle
    lib = ib.capi.lib()
    lib.ibvh_profile_enable(handle, 1)
    prof_steps = 5
    for _ in range(prof_steps):
        step_device()
    finish_pending()
    torch.cuda.synchronize()
    rows = profile_rows(ib, handle)
    lib.ibvh_profile_enable(handle, 0)
    per_kernel, launches = {}, {}
    for k, v in rows:
        per_kernel[k] = per_kernel.get(k, 0.0) + v / prof_steps
        launches[k] = launches.get(k, 0) + 1
    launches = {k: v // prof_steps for k, v in launches.items()}
    gpu_launches = int(sum(launches.values()))
    dom = max(per_kernel, key=per_kernel.get)
    nq = qe - qb
    my_contacts = state["tr"].num_contacts
    # Algorithmic bytes of the dominant kernel per step. Traversal kernels (T1): the tree is read
    # once per traversal pass + the contact output: passes * (Lb + Nb) * N + 2 * Ib * C (ordered = 2 passes,
    # unordered = 1). Build kernels: S2 sort 68 N, S3+S4 gather+merge 76 N, S1 bounds+encode 36 N (wrap path).
    Lb, Nb, Ib = 24, 24, 4
    n_trav = launches.get(dom, 1)
    trav_kernels = ("pyr_leaf_tile_kernel", "pyr_refine_kernel", "tile_flat_kernel", "tile_kernel", "group_walk_kernel", "lvt_packet_kernel", "lvt_thread_kernel")
    if dom in trav_kernels:
        passes = 2 if ordered else 1
        alg_bytes = passes * (Lb + Nb) * n * (nq / n if world > 1 else 1.0) + 2 * Ib * my_contacts
    else:
        alg_bytes = {"onesweep_kernel": 68, "gather_merge_kernel": 76}.get(dom, 36) * n
    dom_ms = per_kernel[dom]
    peak, peak_src = peaks()
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get(dom)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom, "launches_per_step": n_trav, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_step": alg_bytes,
                "kernel_ms_per_step": dom_ms, "share_of_step": dom_ms / ms_step,
                "per_kernel_ms": {k: round(v, 4) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])},
                "traversal_ms_per_step": round(sum(v for k, v in per_kernel.items() if k.startswith(("pyr_", "tile_", "group_walk", "lvt_", "scan_r", "scan_b", "scan_a"))), 4),
                "build_ms_per_step": round(sum(v for k, v in per_kernel.items() if k in ("init_build_kernel", "bounds_kernel", "encode_kernel", "scan_hist_kernel", "onesweep_kernel", "gather_kernel", "gather_merge_kernel", "merge_levels_kernel")), 4)}
    # the same arithmetic for the build and for the traversal as a whole (SURVEY.md §8d byte formulas; 32-bit type set)
    def group(ms, nbytes):
        gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"ms_per_step": ms, "algorithmic_bytes_per_step": nbytes, "achieved": gbs, "frac": gbs / peak, "unit": "GB/s"}
    stage_bytes = {"bounds_kernel": 16 * n, "encode_kernel": 20 * n, "onesweep_kernel": 68 * n, "scan_hist_kernel": 0,
                   "gather_kernel": 48 * n, "gather_merge_kernel": 48 * n}
    shard_frac = (nq / n) if world > 1 else 1.0
    roofline["groups"] = {
        "build (S1+S2+S3+S4 = 196 N)": group(roofline["build_ms_per_step"], 196 * n),
        "traversal (T1 = 48 N + 8 C, this rank's shard)": group(roofline["traversal_ms_per_step"], (Lb + Nb) * n * shard_frac + 2 * Ib * my_contacts),
        "per_build_kernel": {k: group(per_kernel[k], b) for k, b in stage_bytes.items() if k in per_kernel and b > 0},
    }
    # intersection tests of one traversal (this rank's shard), derived by the library from its pair-list sizes
    import ctypes as C
    st4 = (C.c_int64 * 4)()
    lib.ibvh_last_traversal_stats(handle, st4)
    trav_ms = roofline["traversal_ms_per_step"]
    if st4[3] > 0 and trav_ms > 0:
        roofline["traversal_tests"] = {"box_tests_per_step": int(st4[0]), "leaf_tests_per_step": int(st4[1]), "candidate_group_pairs": int(st4[2]),
                                       "intersection_tests_per_s": (int(st4[0]) + int(st4[1])) / (trav_ms * 1e-3), "scope": "this rank's query shard"}

    # ---- e2e: host (pinned) volumes in, contacts back to pinned host memory ------------------------
    host_vols = synth.random_spheres_np(n, seed=SEED) if rank == 0 or world > 1 else None
    pinned_in = ib.DeviceArray.from_numpy(host_vols, pin=True)
    cap = int(ncontacts * 1.05) + 1024
    pinned_out = torch.empty(cap * 8, dtype=torch.uint8).pin_memory()
    e2e_state = {"bvh": state["bvh"], "tr": state["tr"]}

    def step_e2e():
        d_in = pinned_in.to(dev)                                     # H2D of the step's input
        if world == 1:
            bvh = ib.BVH(d_in, ib.BBox(), cache=e2e_state["bvh"])
            tr = ib.traverse(bvh, cache=e2e_state["tr"], ordered=ordered)
            e2e_state["bvh"], e2e_state["tr"] = bvh, tr
            nb = tr.num_contacts * 8
            pinned_out[:nb].copy_(tr.cache1.tensor[:nb], non_blocking=True)   # D2H of the step's result
            torch.cuda.current_stream().synchronize()
            return tr.num_contacts
        if args.build_mode == "replicate" or rank == 0:
            bvh = ib.BVH(d_in, ib.BBox(), cache=e2e_state["bvh"])
            e2e_state["bvh"] = bvh
        bvh = e2e_state["bvh"]
        if args.build_mode == "broadcast":
            dist.broadcast(bvh.leaves.tensor, src=0)
            dist.broadcast(bvh.nodes.tensor, src=0)
        tr = ib.traverse(bvh, cache=e2e_state["tr"], ordered=ordered, query_range=(qb, qe - qb))
        e2e_state["tr"] = tr
        if state["peer"] is not None:
            full, total, _ = state["peer"].gather(tr.cache1.tensor, tr.num_contacts)
        else:
            full, counts = ibdist.gather_shards(tr.cache1.tensor, tr.num_contacts, 8)
            total = int(sum(counts))
        nb = total * 8
        pinned_out[:nb].copy_(full[:nb], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return total

    even = world > 1 and n % world == 0
    shard_lo, shard_hi = (qb * 16, qe * 16) if even else (0, n * 16)      # this rank's slice of the host volumes (bytes)

    def run_e2e_pipelined(k_steps):
        """The same per-step work (H2D of the step's volumes -> BVH -> traverse -> D2H of the step's contact list),
        software-pipelined over three streams with double buffers so that the copies of neighbouring steps overlap
        the kernels (PCIe is full duplex). Every step's input still comes from pinned host memory and every step's
        contacts still land in pinned host memory inside the timed region.
        N > 1: each rank's host holds 1/N of the volumes and receives its own shard of the contact list; the
        volumes are all-gathered over NVLink (NCCL, in place), the build is replicated, the traversal sharded."""
        main = torch.cuda.current_stream(dev)
        if "bufs" not in e2e_state:
            # streams and double buffers are set up ONCE (page-locking a 330 MB host buffer takes tens of milliseconds: round 1
            # did it inside this function, i.e. inside the timed region, which cost every step several milliseconds)
            e2e_state["bufs"] = {
                "s_in": torch.cuda.Stream(dev), "s_out": torch.cuda.Stream(dev),
                "d_in": [ib.DeviceArray.empty(n, pinned_in.dtype, dev) for _ in range(2)],
                "caches": [ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(cap, ib.pair_dtype(), dev), ib.DeviceArray.empty(n, np.int32, dev)) for _ in range(2)],
                "outs": [pinned_out, torch.empty(cap * 8, dtype=torch.uint8).pin_memory()],
            }
        B = e2e_state["bufs"]
        s_in, s_out, d_in, caches, outs = B["s_in"], B["s_out"], B["d_in"], B["caches"], B["outs"]
        in_ready = [torch.cuda.Event() for _ in range(2)]
        comp_done = [torch.cuda.Event() for _ in range(2)]
        out_done = [torch.cuda.Event() for _ in range(2)]
        bvh_prev = e2e_state["bvh"]
        use_fused = world > 1 and fused and state["peer"] is not None
        skip = set(os.environ.get("IBVH_E2E_SKIP", "").split(","))        # diagnosis only: h2d, allgather, d2h
        if use_fused and e2e_state.get("peers") is None:
            # two symmetric buffers: step k's slice of the gathered list leaves for the host while step k+1 writes the other
            e2e_state["peers"] = [state["peer"], ibdist.PeerGather(state["peer"].capacity_bytes // 8, 8, dev)]
        copied = {"pairs": 0}

        def h2d(k):
            b = k % 2
            with torch.cuda.stream(s_in):
                s_in.wait_event(comp_done[b])                 # step k-2 no longer reads d_in[b]
                if "h2d" not in skip:
                    d_in[b].tensor[shard_lo:shard_hi].copy_(pinned_in.tensor[shard_lo:shard_hi], non_blocking=True)
                in_ready[b].record(s_in)

        for ev in comp_done + out_done:
            ev.record(main)
        h2d(0)
        last = 0
        for k in range(k_steps):
            b = k % 2
            if k + 1 < k_steps:
                h2d(k + 1)                                     # next step's input travels while this step computes
            main.wait_event(in_ready[b])
            main.wait_event(out_done[b])                       # step k-2's contacts have left caches[b]
            if even and "allgather" not in skip:
                dist.all_gather_into_tensor(d_in[b].tensor, d_in[b].tensor[shard_lo:shard_hi])
            bvh = ib.BVH(d_in[b], ib.BBox(), cache=bvh_prev)
            if world == 1:
                tr = ib.traverse(bvh, cache=caches[b], ordered=ordered)
                lo, hi = 0, tr.num_contacts
                src_t = tr.cache1.tensor
                caches[b] = tr
            elif use_fused:
                # the traversal kernel itself writes every contact into every rank's list (NVLink multicast); each rank's host
                # then receives ITS 1/N slice of the gathered list, so the hosts together receive the whole list once
                tr = ib.traverse(bvh, ordered=False, query_range=(qb, qe - qb), peer=e2e_state["peers"][b])
                tot = tr.num_contacts
                lo, hi = tot * rank // world, tot * (rank + 1) // world
                src_t = e2e_state["peers"][b].list_area()
            else:
                tr = ib.traverse(bvh, cache=caches[b], ordered=ordered, query_range=(qb, qe - qb))
                lo, hi = 0, tr.num_contacts
                src_t = tr.cache1.tensor
                caches[b] = tr
            bvh_prev = bvh
            comp_done[b].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(comp_done[b])
                if "d2h" not in skip:
                    outs[b][:(hi - lo) * 8].copy_(src_t[lo * 8:hi * 8], non_blocking=True)
                out_done[b].record(s_out)
            last = hi - lo
        s_out.synchronize(); s_in.synchronize(); main.synchronize()
        return last

    pipelined = world == 1 or args.build_mode == "replicate"
    if pipelined:
        run_e2e_pipelined(3)
        sync_all()
        t0 = time.perf_counter()
        e0.record()
        nc = run_e2e_pipelined(args.steps)
        e1.record()
        sync_all()
        e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / args.steps
    # the strictly sequential figure (no overlap between steps; N > 1: every rank uploads all volumes and the
    # gathered contact list comes back), for reference
    seq_steps = max(3, args.steps // 4) if pipelined else args.steps
    for _ in range(3):
        step_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(seq_steps):
        nc_seq = step_e2e()
    sync_all()
    e2e_seq_ms = (time.perf_counter() - t0) * 1e3 / seq_steps
    if not pipelined:
        e2e_ms, nc = e2e_seq_ms, nc_seq
    h2d_bytes = n * 16 if (world == 1 or (pipelined and even)) else n * 16 * world
    if world > 1:
        t = torch.tensor([e2e_ms, e2e_seq_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms, e2e_seq_ms = float(t[0].item()), float(t[1].item())
        t = torch.tensor([nc], dtype=torch.int64, device=dev)
        if pipelined:
            dist.all_reduce(t)                      # contacts copied back, summed over the ranks' shards
        nc = int(t.item())
    e2e = {"value": n / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(nc * 8),
           "ms_per_step": e2e_ms, "sequential_ms_per_step": e2e_seq_ms,
           "note": "public API (BVH + traverse) from pinned host volumes to pinned host contacts; copies of neighbouring steps overlap "
                   "the kernels (3 streams, double buffers); N > 1: each rank's host holds 1/N of the volumes (all-gathered over NVLink "
                   "after the upload) and receives its own shard of the contact list; sequential_ms_per_step is the same without any "
                   "overlap (N > 1: every rank uploads all volumes, rank-gathered contact list comes back)"}

    # ---- secondary metric of BASELINE.json: rays/s @ 1 M leaves (configs[3]) ------------------------------
    # 1000 x 1000 mesh-like shell of spheres, R random rays (origins U[-1.5,1.5)^3, directions uniform on S^2),
    # rays sharded by contiguous ranges over the ranks; every rank builds the (1 M-leaf, 0.2 ms) tree itself.
    rays = None
    if not args.no_rays:
        R = int(os.environ.get("IBVH_BENCH_RAYS", 100_000_000))
        rb = ibdist.shard_bounds(R, world)[rank]
        nr = rb[1] - rb[0]
        shell = synth.shell_spheres_np(1000, 1000)
        rbvh = ib.BVH(shell, ib.BBox(), device=dev)
        rp, rd = synth.random_rays_torch(nr, dev, seed=7, start=rb[0])
        rt = ib.traverse_rays(rbvh, rp, rd, ordered=ordered, id_base=rb[0])
        rcache = ib.BVHTraversal(1, 0, 0, 0, ib.DeviceArray.empty(int(rt.num_contacts * 1.02) + 1024, ib.pair_dtype(), dev), rt.cache2)
        ray_peer = None
        if world > 1 and args.gather != "nccl":
            tot = torch.tensor([rt.num_contacts], dtype=torch.int64, device=dev)
            dist.all_reduce(tot)
            try:
                ray_peer = ibdist.PeerGather(int(int(tot.item()) * 1.10) + 8192 * world, 8, dev)   # hit list of ALL rays on every rank
            except Exception:
                ray_peer = None

        rays_fused = ray_peer is not None and args.gather == "fused" and not ordered and bool(ray_peer.peer.multicast)

        def step_rays():
            if rays_fused:          # the rays kernel writes every hit into every rank's list (multimem.st over NVLink)
                return ib.traverse_rays(rbvh, rp, rd, ordered=False, id_base=rb[0], peer=ray_peer).num_contacts
            t = ib.traverse_rays(rbvh, rp, rd, cache=rcache, ordered=ordered, id_base=rb[0])
            if world > 1:
                if ray_peer is not None:
                    _, total_hits, _ = ray_peer.gather(t.cache1.tensor, t.num_contacts)     # ibvh_allgather_pairs: rank order = ray order
                    return total_hits
                _, cs = ibdist.gather_shards(t.cache1.tensor, t.num_contacts, 8)
                return int(sum(cs))
            return t.num_contacts

        def step_rays_sharded():
            """Every rank traces its ray range; the hits stay where they are found (rank order = ray order: the shards
            concatenated are the single-GPU list). No exchange."""
            t = ib.traverse_rays(rbvh, rp, rd, cache=rcache, ordered=ordered, id_base=rb[0])
            return t

        def timed_rays(fn, steps=3, warm_steps=2):
            for _ in range(warm_steps):
                out = fn()
            sync_all()
            e0.record()
            for _ in range(steps):
                out = fn()
            e1.record()
            sync_all()
            ms = e0.elapsed_time(e1) / steps
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms, out

        ray_verify = None
        gathered = None
        if world == 1:
            rms, hits = timed_rays(step_rays)
        else:
            # headline: rays sharded by range, hits left sharded (BASELINE configs[3]: "rays sharded across 1/2/4/8 B200")
            rms, tsh = timed_rays(step_rays_sharded)
            mine = pair_checksum(torch, tsh.cache1.tensor, tsh.num_contacts)
            acc = torch.tensor([mine[0], mine[1], mine[2]], dtype=torch.int64, device=dev)
            dist.all_reduce(acc)                                         # (int64 sums wrap mod 2^64, like the checksums)
            hits = int(acc[0].item())
            # the same with every rank receiving ALL hits (fused into the rays kernel / library all-gather / NCCL)
            gms, ghits = timed_rays(step_rays)
            gathered = {"ms_per_step": gms, "value": R / (gms * 1e-3), "unit": "rays/s", "hits_per_step": int(ghits),
                        "how": ("fused into the rays kernel: multimem.st over NVLink into every rank's list" if rays_fused else
                                "ibvh_allgather_pairs over NVLink peer memory" if args.gather != "nccl" else "NCCL")}
            # verification on rank 0: the whole ray set on one GPU against the sum of the shards (count + checksum)
            ok = 1
            if rank == 0:
                fp, fd = synth.random_rays_torch(R, dev, seed=7)
                full = ib.traverse_rays(rbvh, fp, fd, ordered=False)
                want = pair_checksum(torch, full.cache1.tensor, full.num_contacts)
                wrap = lambda v: (v + (1 << 63)) % (1 << 64) - (1 << 63)
                got = (int(acc[0].item()), int(acc[1].item()), int(acc[2].item()))
                ok = 1 if (want[0] == got[0] and wrap(want[1]) == got[1] and wrap(want[2]) == got[2] and int(ghits) == want[0]) else 0
                del fp, fd, full
            okt = torch.tensor([ok], dtype=torch.int64, device=dev)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            ray_verify = {"shards_equal_single_gpu_hit_list": bool(okt.item()),
                          "what": "hit count + order-independent checksum summed over the ranks' shards against rank 0 tracing all rays alone; the gathered variant's total too"}
        rays = {"metric": "rays/s @1M leaves", "value": R / (rms * 1e-3), "unit": "rays/s", "ms_per_step": rms, "rays": R, "hits_per_step": int(hits if world > 1 else hits),
                "workload": "configs[3]: 1000x1000 shell of BSphere{Float32} (1 M leaves, BBox{Float32} nodes), %d random rays, traverse_rays (LVT), "
                            "rays sharded by contiguous ranges over %d GPU(s)%s" % (R, world, "" if world == 1 else "; every rank keeps the hits of its own rays (rank order = ray order); `gathered_to_every_rank` is the same step with all hits delivered to every rank"),
                "gathered_to_every_rank": gathered, "verify": ray_verify, "scaling": "strong"}
        del rp, rd, rcache, rt


    # ---- the other BASELINE.json workloads (short timed runs, reported under `workloads`) ------------------------
    workloads = {}

    def rank_max(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def shard_gather(tr, peer_obj):
        """N > 1: all-gather this rank's shard (library kernel over NVLink peer memory, else NCCL). Returns the gathered total."""
        if world == 1:
            return tr.num_contacts
        if peer_obj is not None:
            return peer_obj.gather(tr.cache1.tensor, tr.num_contacts)[1]
        return int(sum(ibdist.gather_shards(tr.cache1.tensor, tr.num_contacts, tr.cache1.dtype.itemsize)[1]))

    if "ordered" in extra and not ordered:
        # the headline step in the reference's own contact order (ascending query leaf, DFS order): the drop-in default
        st_o = {"bvh": state["bvh"], "tr": None}

        def step_ordered():
            bvh = ib.BVH(src, ib.BBox(), cache=st_o["bvh"])
            tr = ib.traverse(bvh, cache=st_o["tr"], ordered=True, query_range=(qb, qe - qb) if world > 1 else None)
            st_o["bvh"], st_o["tr"] = bvh, tr
            return shard_gather(tr, state["peer"])

        ms_o, c_o = timed_steps(torch, step_ordered, 5, 3)
        ms_o = rank_max(ms_o)
        workloads["ordered"] = {"metric": METRIC, "value": n / (ms_o * 1e-3), "unit": UNIT, "ms_per_step": ms_o, "contacts_per_step": int(c_o),
                                "workload": "the headline step with contacts in the reference's order (count -> scan -> write protocol; "
                                            "byte-identical to the oracle's list); N > 1: ordered shards all-gathered in rank order"}
        st_o.clear()

    if "bfs" in extra and world == 1:
        # the headline step with the reference's other traversal algorithm, BFSTraversal (src/traverse/breadth_first): level-synchronous
        # BVTT descent from the default start level (levels / 2); same contact set (checked here by count + order-independent
        # checksum against an LVT traversal of the same tree), exact num_checks
        st_b = {"bvh": state["bvh"], "tr": None}

        def step_bfs():
            bvh = ib.BVH(src, ib.BBox(), cache=st_b["bvh"])
            tr = ib.traverse(bvh, ib.BFSTraversal(), cache=st_b["tr"])
            st_b["bvh"], st_b["tr"] = bvh, tr
            return tr

        ms_b, tr_b = timed_steps(torch, step_bfs, 5, 3)
        lvt = ib.traverse(st_b["bvh"], ordered=False)
        same = pair_checksum(torch, tr_b.cache1.tensor, tr_b.num_contacts) == pair_checksum(torch, lvt.cache1.tensor, lvt.num_contacts)
        lib.ibvh_profile_enable(handle, 1)
        step_bfs()
        torch.cuda.synchronize()
        prow = profile_rows(ib, handle)
        lib.ibvh_profile_enable(handle, 0)
        pk = {}
        for k, v in prow:
            pk[k] = pk.get(k, 0.0) + v
        trav_ms = sum(v for k, v in pk.items() if k.startswith("bfs_"))
        workloads["bfs"] = {"metric": METRIC, "value": n / (ms_b * 1e-3), "unit": UNIT, "ms_per_step": ms_b, "contacts_per_step": int(tr_b.num_contacts),
                            "num_checks": int(tr_b.num_checks), "start_level": int(tr_b.start_level1), "traverse_kernels_ms": round(trav_ms, 4),
                            "checks_per_s": tr_b.num_checks / (trav_ms * 1e-3) if trav_ms > 0 else None,
                            "same_contact_set_as_lvt": bool(same),
                            "per_kernel_ms": {k: round(v, 4) for k, v in sorted(pk.items(), key=lambda kv: -kv[1])},
                            "workload": "the headline step with traverse(bvh, BFSTraversal()) — the reference's breadth-first algorithm (fewest checks, "
                                        "BVTT lists in library scratch); contacts unordered like the reference's GPU backend"}
        del lvt, tr_b
        st_b.clear()

    if "reference_shaped" in extra and world == 1:
        # proxy of the reference's CUDA.jl backend (cannot run here: no Julia): the launches BVH(...) + traverse(...) make through
        # AcceleratedKernels, written naively in CUDA — struct-moving merge sort, two mapreduce passes with host read-backs, one
        # merge launch per level, one thread per leaf with a private stack, count -> scan -> write
        st_r = {"tr": None}

        def step_proxy():
            bvh = ib.BVH(src, ib.BBox(), reference_shaped=True)
            tr = ib.traverse(bvh, cache=st_r["tr"], ordered=True, reference_shaped=True)
            st_r["tr"] = tr
            return tr.num_contacts

        ms_r, c_r = timed_steps(torch, step_proxy, 3, 2)
        lib.ibvh_profile_enable(handle, 1)
        step_proxy()
        torch.cuda.synchronize()
        prow = profile_rows(ib, handle)
        lib.ibvh_profile_enable(handle, 0)
        pk = {}
        for k, v in prow:
            pk[k] = pk.get(k, 0.0) + v
        workloads["reference_shaped"] = {
            "metric": METRIC, "value": n / (ms_r * 1e-3), "unit": UNIT, "ms_per_step": ms_r, "contacts_per_step": int(c_r),
            "speedup_of_product_path": ms_r / ms_step,
            "per_kernel_ms": {k: round(v, 4) for k, v in sorted(pk.items(), key=lambda kv: -kv[1])},
            "label": "PROXY, not ImplicitBVH.jl: a naive reference-shaped CUDA build + LVT (ibvh_build_reference_shaped + "
                     "IBVH_TRAVERSE_REFERENCE_SHAPED) standing in for the reference's CUDA.jl backend, which cannot run in this image"}
        st_r.clear()

    if "pair" in extra:
        # configs[2]: BVH-vs-BVH, 5 M + 5 M leaves, the target tree built to levels-13 (about 2^10 roots) and fully;
        # the query leaves (bvh1) are sharded over the ranks, both builds replicated
        np_ = 5_000_000 if n >= 10_000_000 else max(1000, n // 2)
        sc = synth.sphere_radius_scale(np_)
        v1 = synth.random_spheres_torch(np_, dev, seed=42, scale=sc)
        v2 = synth.random_spheres_torch(np_, dev, seed=43, scale=sc)
        d1 = ib.DeviceArray(v1.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
        d2 = ib.DeviceArray(v2.view(torch.uint8).reshape(-1), ib.BSphere().dtype)
        pqb, pqe = ibdist.shard_bounds(np_, world)[rank]
        levels = ib.ImplicitTree(np_).levels
        res = {}
        for name, bl in (("partial_build", max(1, levels - 13)), ("full_build", 1)):
            st_p = {"b1": None, "b2": None, "tr": None}

            def step_pair():
                b1 = ib.BVH(d1, ib.BBox(), cache=st_p["b1"])
                b2 = ib.BVH(d2, ib.BBox(), built_level=bl, cache=st_p["b2"])
                tr = ib.traverse(b1, b2, cache=st_p["tr"], ordered=False, query_range=(pqb, pqe - pqb) if world > 1 else None)
                st_p["b1"], st_p["b2"], st_p["tr"] = b1, b2, tr
                return shard_gather(tr, state["peer"] if world > 1 else None)

            def trav_only():
                tr = ib.traverse(st_p["b1"], st_p["b2"], cache=st_p["tr"], ordered=False, query_range=(pqb, pqe - pqb) if world > 1 else None)
                st_p["tr"] = tr
                return tr.num_contacts

            try:
                ms_p, c_p = timed_steps(torch, step_pair, 5, 3)
                ms_t, _ = timed_steps(torch, trav_only, 5, 1)
            except Exception as ex:                                  # e.g. the peer list area sized for the headline is too small
                res[name] = {"error": repr(ex)[:300]}
                continue
            ms_p, ms_t = rank_max(ms_p), rank_max(ms_t)
            res[name] = {"built_level2": bl, "start_level2": bl, "ms_per_step": ms_p, "traverse_ms": ms_t, "contacts_per_step": int(c_p),
                         "query_leaves_per_s": np_ / (ms_p * 1e-3), "leaves_per_s": 2 * np_ / (ms_p * 1e-3)}
            st_p.clear()
        workloads["pair"] = {"metric": "pair traversal: (5 M + 5 M) leaves/s, two builds + traverse(bvh1, bvh2)", "unit": "leaves/s",
                             "value": res.get("partial_build", {}).get("leaves_per_s"), "leaves": [np_, np_], **res,
                             "workload": "configs[2]: two independent draws (seeds 42, 43) of the config-2 law in one unit cube; step = BVH(bvh1) + "
                                         "BVH(bvh2; built_level) + traverse(bvh1, bvh2) unordered; bvh1's leaves are the queries, sharded by contiguous "
                                         "ranges over %d GPU(s)%s" % (world, "" if world == 1 else ", shards all-gathered to every rank")}
        del v1, v2, d1, d2

    if "rebuild64" in extra:
        # configs[4]: 100 M leaves, UInt64 Morton, Int64 index; each step perturbs the centres in place, rebuilds in place with
        # cache=bvh (BVH(bvh.leaves, cache=bvh)) and finds the contacts with cache=previous traversal
        n64 = int(os.environ.get("IBVH_BENCH_N64", 100_000_000 if n >= 10_000_000 else 10 * n))
        o64 = ib.BVHOptions(index=np.int64, morton=ib.DefaultMortonAlgorithm(np.uint64))
        try:
            v64 = synth.random_spheres_torch(n64, dev, seed=42)
            bvh64 = ib.BVH(ib.DeviceArray(v64.view(torch.uint8).reshape(-1), ib.BSphere().dtype), ib.BBox(), options=o64)
            del v64
            qb64, qe64 = ibdist.shard_bounds(n64, world)[rank]
            F = bvh64.leaves.tensor.view(torch.float32).reshape(n64, 8)
            s64 = synth.sphere_radius_scale(n64)
            st64 = {"bvh": bvh64, "tr": None, "k": 0}
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            acc = {"build": 0.0, "trav": 0.0, "steps": 0}

            def step64(timed):
                st64["k"] += 1
                for c in range(3):                                   # the simulation's move (not part of the measured path)
                    F[:, c] += (synth.uniform_torch(1000 + st64["k"], n64, c, dev) - 0.5) * (s64 / 2)
                ev[0].record()
                b = ib.BVH(st64["bvh"].leaves, ib.BBox(), cache=st64["bvh"], options=o64)
                ev[1].record()
                tr = ib.traverse(b, cache=st64["tr"], ordered=False, query_range=(qb64, qe64 - qb64) if world > 1 else None)
                ev[2].record()
                st64["bvh"], st64["tr"] = b, tr
                torch.cuda.synchronize()
                if timed:
                    acc["build"] += ev[0].elapsed_time(ev[1]); acc["trav"] += ev[1].elapsed_time(ev[2]); acc["steps"] += 1
                return tr.num_contacts

            for _ in range(3):
                step64(False)
            c64 = 0
            for _ in range(4):
                c64 = step64(True)
            bms, tms = rank_max(acc["build"] / acc["steps"]), rank_max(acc["trav"] / acc["steps"])
            ctot = c64
            if world > 1:
                t = torch.tensor([c64], dtype=torch.int64, device=dev)
                dist.all_reduce(t)
                ctot = int(t.item())
            workloads["rebuild64"] = {"metric": "cached rebuild + contact leaves/s @100M, UInt64 / Int64", "unit": UNIT, "leaves": n64,
                                      "value": n64 / ((bms + tms) * 1e-3), "ms_per_step": bms + tms, "build_ms": bms, "traverse_ms": tms,
                                      "contacts_per_step": int(ctot),
                                      "workload": "configs[4]: config-2 law at %d leaves, 32-byte leaves (UInt64 Morton, Int64 index), 16-byte pairs; per step the "
                                                  "centres move by U[-s/4, s/4), then BVH(bvh.leaves, BBox; cache=bvh) in place + traverse(cache=prev) unordered; "
                                                  "the (replicated) rebuild and the query-range sharded traversal over %d GPU(s) are timed with CUDA events, the "
                                                  "move is not; N > 1: every rank keeps its own shard (no gather)" % (n64, world)}
            del F, bvh64
            st64.clear()
        except Exception as ex:
            workloads["rebuild64"] = {"error": repr(ex)[:300]}
        torch.cuda.empty_cache()
        lib.ibvh_release_workspace(handle)

    clocks = sampler.stop() if rank == 0 else None      # sampled over the timed, e2e, ray and extra-workload regions

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload ----------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        ns = cpu_calibrated_size(threads, 12.0)
        O, np_, leaves, nodes, contacts, counts = cpu_setup(ns)
        dt, tot = cpu_step(O, np_, leaves, nodes, contacts, counts, threads)
        cpu_baseline = {"value": ns / dt, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"one build + two-pass LVT traversal of a {ns}-leaf instance of the same law ({tot} contacts) in {dt:.2f} s; "
                                  f"C++17 restatement of the reference algorithm (oracle/), {threads} threads"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": bench_config(),
            "details": {"contacts_per_step": int(ncontacts),
                       "contact_order": "reference (ascending query, DFS order; count+scan+write)" if ordered else "unordered (one pass, buffered warp-aggregated atomics; identical as a sorted list)",
                       "parallelism": "single GPU" if world == 1 else (("build replicated on every rank (deterministic, bit-identical)" if args.build_mode == "replicate" else "build on rank 0 + NCCL broadcast of the tree") + f", query-range sharded traversal over {world} GPUs, " + ("traversal fused with the all-gather: contacts written into every rank's list by the traversal kernel (multimem.st over NVLink)" if fused else "contact shards all-gathered by the library's peer-memory kernel (NVLink multicast stores)" if args.gather in ("peer", "fused") else "NCCL all-gather of the contact shards")),
                       "host_sync": ("deferred traversal (IBVH_TRAVERSE_DEFER): the next build is enqueued before the host reads a step's contact count; every count is read inside the timed region"
                                     if (world == 1 and not ordered and not args.no_defer) else "synchronous calls: the host reads each step's contact count before the next build is enqueued"),
                       "l2_policy": "inputs larger than L2 (160 MB volumes + 240 MB leaves + 240 MB nodes per step vs 126 MB L2)"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": gpu_launches, "clocks": clocks,
            "secondary": rays, "workloads": workloads, "verify": verify,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
