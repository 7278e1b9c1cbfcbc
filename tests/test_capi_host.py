"""The C-ABI library loads, exports every symbol include/ibvh.h declares, and its host-only integer
math agrees with the oracle and the reference's known answers. No GPU needed (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "ibvh.h")).read()
    return sorted(set(re.findall(r"IBVH_API\s+[\w\s\*]+?\b(ibvh_\w+)\s*\(", txt)))


def test_every_header_symbol_is_exported_and_bound(ib):
    lib = ib.capi.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ibvh.h but not exported"
        assert s in ib.capi.SIGNATURES, f"{s} has no ctypes signature"
    assert sorted(ib.capi.SIGNATURES) == syms


def test_version_and_status_strings(ib):
    lib = ib.capi.lib()
    assert lib.ibvh_version() == 100
    assert ib.capi.status_string(1) == "ArgumentError"
    assert ib.capi.status_string(2) == "DomainError"


def test_tree_known_answers_through_capi(ib, golden):
    for case in golden["implicit_tree"]:
        t = ib.ImplicitTree(case["n"])
        for k in ("levels", "real_leaves", "virtual_leaves", "real_nodes", "virtual_nodes"):
            assert getattr(t, k) == case[k]
        for idx, want in case["memory_index"]:
            assert ib.memory_index(t, idx) == want
        for lvl, want in case["level_indices"]:
            assert list(ib.level_indices(t, lvl)) == want
        for idx, want in case["isvirtual"]:
            assert ib.isvirtual(t, idx) == want


def test_tree_matches_oracle_sweep(ib, O):
    for n in list(range(1, 300)) + [1000, 4097, 10**5, 10**6 + 7, 10**7, 10**8, 2**31 - 5]:
        t = ib.ImplicitTree(n)
        o = O.tree_shape(n)
        assert (t.levels, t.real_nodes, t.virtual_leaves, t.virtual_nodes) == (o["levels"], o["real_nodes"], o["virtual_leaves"], o["virtual_nodes"])
        assert (t.skips() == o["skips"]).all()
        assert ib.capi.lib().ibvh_num_nodes(n) == o["real_nodes"] - o["real_leaves"]
        for idx in {1, 2 ** (t.levels - 1), 2 ** t.levels - 1, max(1, 2 ** (t.levels - 1) + n - 1)}:
            assert ib.isvirtual(t, idx) == O.isvirtual(n, idx)
            if not O.isvirtual(n, idx):
                assert ib.memory_index(t, idx) == O.memory_index(n, idx)


def test_tree_shape_table_from_survey(ib):
    """tree shapes at the benchmark configs."""
    for n, levels, nodes in ((100_000, 18, 100_006), (1_000_000, 21, 1_000_007), (5_000_000, 24, 5_000_009),
                             (10_000_000, 25, 10_000_009), (100_000_000, 28, 100_000_007)):
        t = ib.ImplicitTree(n)
        assert t.levels == levels and t.real_nodes - t.real_leaves == nodes
    assert ib.ImplicitTree(10_000_000).virtual_leaves == 6_777_216


def test_domain_and_bounds_errors(ib):
    with pytest.raises(ib.DomainError):
        ib.ImplicitTree(0)
    t = ib.ImplicitTree(5)
    with pytest.raises(IndexError):
        ib.memory_index(t, 16)
    with pytest.raises(IndexError):
        ib.level_indices(t, 5)


def test_build_level_rule(ib, golden):
    lib = ib.capi.lib()
    g = golden["build_level"]
    out = C.c_int64()
    for c in g["cases"]:
        if isinstance(c["built_level"], float):
            assert lib.ibvh_compute_build_level(g["levels"], 1, 0, c["built_level"], C.byref(out)) == 0
        else:
            assert lib.ibvh_compute_build_level(g["levels"], 0, c["built_level"], 0.0, C.byref(out)) == 0
        assert out.value == c["expect"]
    assert lib.ibvh_compute_build_level(8, 0, 0, 0.0, C.byref(out)) == ib.capi.ERR_ARGUMENT
    assert lib.ibvh_compute_build_level(8, 0, 9, 0.0, C.byref(out)) == ib.capi.ERR_ARGUMENT
    assert lib.ibvh_compute_build_level(8, 1, 0, 1.5, C.byref(out)) == ib.capi.ERR_ARGUMENT


def test_leaf_layouts_match_oracle(ib, O):
    for kind in (0, 1):
        for ib_ in (4, 8):
            for mb in (2, 4, 8):
                vol = ib.VolumeType(kind, 4)
                dt = ib.leaf_dtype(vol, {4: np.int32, 8: np.int64}[ib_], {2: np.uint16, 4: np.uint32, 8: np.uint64}[mb])
                t = ib.capi.Types(kind, 4, ib_, mb, 1, 0)
                assert dt.itemsize == ib.capi.lib().ibvh_leaf_bytes(C.byref(t)) == O.leaf_dtype(kind, 4, ib_, mb).itemsize
                assert dt == O.leaf_dtype(kind, 4, ib_, mb)
    assert ib.leaf_dtype(ib.BSphere()).itemsize == 24 and ib.leaf_dtype(ib.BSphere(), np.int64, np.uint64).itemsize == 32
    assert ib.pair_dtype(np.int32).itemsize == 8 and ib.pair_dtype(np.int64).itemsize == 16


def test_options_validation(ib):
    ib.BVHOptions()
    for bad in ("num_threads", "min_mortons_per_thread", "min_sorts_per_thread", "min_boundings_per_thread",
                "min_traversals_per_thread", "block_size"):
        with pytest.raises(ib.ArgumentError):
            ib.BVHOptions(**{bad: 0})
    with pytest.raises(ib.ArgumentError):
        ib.DefaultMortonAlgorithm(np.uint8)
    o = ib.BVHOptions(index=np.int64, morton=ib.DefaultMortonAlgorithm(np.uint64))
    assert o.index_dtype == np.int64 and o.morton.eltype == np.uint64


def test_no_cpu_fallback(ib):
    """Without a CUDA device the product path must fail loudly, never fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ib.CudaError):
        ib.BVH(ib.bspheres([[0, 0, 0], [1, 0, 0]], [1.0, 1.0]))
    out = C.c_void_p()
    assert ib.capi.lib().ibvh_create(C.byref(out), 0) == ib.capi.ERR_CUDA


def test_bfs_default_start_level_and_null_arguments(ib, O):
    """default_start_level(bvh, ::BFSTraversal) = max(levels / 2, built_level) (breadth_first/breadth_first.jl:4-6), host-only;
    the BFS / sort entry points reject null handles and arguments with a status, not a crash."""
    lib = ib.capi.lib()
    for n in (1, 2, 5, 11, 1000, 10_000_000, 100_000_000):
        levels = ib.ImplicitTree(n).levels
        for built in (1, 2, levels // 2, levels):
            want = max(levels // 2, built)
            assert lib.ibvh_bfs_default_start_level(levels, built) == want
            assert O.bfs_default_start_level(n, built) == want
    total, checks = C.c_int64(7), C.c_int64(7)
    assert lib.ibvh_traverse_bfs_single(None, None, None, None, 0, C.byref(total), C.byref(checks), None) == ib.capi.ERR_ARGUMENT
    assert lib.ibvh_traverse_bfs_pair(None, None, None, 1, 1, 0, None, 0, C.byref(total), C.byref(checks), None) == ib.capi.ERR_ARGUMENT
    assert lib.ibvh_traverse_bfs_rays(None, None, None, None, 0, None, None, 0, C.byref(total), C.byref(checks), None) == ib.capi.ERR_ARGUMENT
    assert lib.ibvh_sort_contacts(None, None, 0, 4, 0, None, None) == ib.capi.ERR_ARGUMENT


def test_julia_extension_ccalls_match_the_header():
    """ext/ImplicitBVHB200Ext.jl cannot be executed here (no Julia): at least every `ccall` in it must name an entry point
    of include/ibvh.h with the header's number of arguments, and its struct mirrors must have the C structs' field counts."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "ibvh.h")).read(), flags=re.S)
    decl = {}
    for m in re.finditer(r"IBVH_API\s+[\w\s\*]+?\b(ibvh_\w+)\s*\(([^;]*?)\)\s*;", header, flags=re.S):
        params = m.group(2).strip()
        decl[m.group(1)] = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
    jl = open(os.path.join(root, "ext", "ImplicitBVHB200Ext.jl")).read()

    def top_level_items(text):
        depth, items, cur = 0, [], ""
        for ch in text:
            depth += ch == "{"
            depth -= ch == "}"
            if ch == "," and depth == 0:
                items.append(cur.strip()); cur = ""
            else:
                cur += ch
        items.append(cur.strip())
        return [i for i in items if i]

    calls = list(re.finditer(r"ccall\(\(:(\w+),\s*LIB\),\s*[\w\{\}]+,\s*\(([^)]*)\)", jl, flags=re.S))
    assert len(calls) >= 16
    for m in calls:
        name, types = m.group(1), top_level_items(m.group(2))
        assert name in decl, f"{name}: not declared in include/ibvh.h"
        assert decl[name] == len(types), f"{name}: header has {decl[name]} arguments, the ccall passes {len(types)}"
    for needed in ("ibvh_build", "ibvh_traverse_single", "ibvh_traverse_pair", "ibvh_traverse_rays", "ibvh_traverse_bfs_single",
                   "ibvh_traverse_bfs_pair", "ibvh_traverse_bfs_rays", "ibvh_sort_contacts", "ibvh_traverse_finish", "ibvh_traverse_cancel"):
        assert any(m.group(1) == needed for m in calls), f"the extension does not bind {needed}"
    # struct mirrors: field counts of ibvh_types_t / ibvh_bvh_t / ibvh_traverse_params_t
    def c_fields(struct):
        body = re.search(r"typedef struct " + struct + r"\s*\{(.*?)\}", header, flags=re.S).group(1)
        return sum(len(stmt.split(",")) for stmt in body.split(";") if stmt.strip())
    def jl_fields(struct):
        body = re.search(r"struct " + struct + r"\b[^\n]*\n(.*?)\nend", jl, flags=re.S).group(1)
        return sum(len([f for f in line.split("#")[0].split(";") if f.strip()]) for line in body.splitlines())
    assert c_fields("ibvh_types") == jl_fields("Types") == 6
    assert c_fields("ibvh_bvh") == jl_fields("CBvh") == 6
    assert c_fields("ibvh_traverse_params") == jl_fields("Params") == 7


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "implicitbvh.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "synth.py" and "oracle" in txt.lower() and "import oracle" not in txt, f


def test_ctypes_structs_match_the_header_layout(ib, tmp_path):
    """The ctypes mirrors of the ABI structs have the sizes and field offsets the C compiler gives include/ibvh.h."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    capi = ib.capi
    structs = {"ibvh_types_t": capi.Types, "ibvh_tree_t": capi.Tree, "ibvh_bvh_t": capi.Bvh,
               "ibvh_traverse_params_t": capi.TraverseParams, "ibvh_peer_t": capi.Peer}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "ibvh.h")}"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = {}
    for ln in out.strip().splitlines():
        cname, key, val = ln.split()
        got[(cname, key)] = int(val)
    for cname, cls in structs.items():
        assert got[(cname, "size")] == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
    # status codes and flags
    hdr = open(os.path.join(ROOT, "include", "ibvh.h")).read()
    for name, val in (("IBVH_ERR_CAPACITY", capi.ERR_CAPACITY), ("IBVH_ERR_PEER", capi.ERR_PEER), ("IBVH_ERR_AGAIN", capi.ERR_AGAIN)):
        assert re.search(rf"{name}\s*=\s*{val}\b", hdr), name
    for name, val in (("IBVH_TRAVERSE_UNORDERED", capi.TRAVERSE_UNORDERED), ("IBVH_TRAVERSE_COUNTS_VALID", capi.TRAVERSE_COUNTS_VALID),
                      ("IBVH_TRAVERSE_DEFER", capi.TRAVERSE_DEFER), ("IBVH_TRAVERSE_WALK", capi.TRAVERSE_WALK), ("IBVH_MAX_PEERS", capi.MAX_PEERS)):
        assert re.search(rf"#define\s+{name}\s+{val}u?\b", hdr), name


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU oracle timed as the reference arm) prints ONE JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, IBVH_BENCH_N="200000")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "leaves/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"]


def test_fused_gather_regions_and_gap_filling(ib):
    """Host logic of the fused multi-GPU traversal (no GPU): the regions PeerGather.set_regions lays out and the moves
    ibvh_peer_compact_plan derives close every gap — applying them to a segmented list leaves the entries [0, total)
    a permutation of all ranks' entries."""
    import ctypes as C
    lib = ib.capi.lib()
    rng = np.random.default_rng(3)
    for world in (1, 2, 3, 4, 8, 16):
        for trial in range(40):
            counts = rng.integers(0, 5000, world).astype(np.int64)
            if trial % 5 == 0:
                counts[rng.integers(0, world)] = 0
            slack = rng.integers(0, 400, world).astype(np.int64) * 2
            begin = np.zeros(world + 1, np.int64)
            begin[1:] = np.cumsum(((counts + 1) & ~1) + slack)
            total = int(counts.sum())
            lst = np.full(int(begin[-1]) + 8, -1, np.int64)
            for r in range(world):
                lst[begin[r]:begin[r] + counts[r]] = r * 1_000_000 + np.arange(counts[r])
            want = np.sort(lst[lst >= 0])
            src, dst, ln = (np.zeros(64, np.int64) for _ in range(3))
            p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
            nm = lib.ibvh_peer_compact_plan(world, p(begin), p(counts), p(src), p(dst), p(ln), 64)
            assert 0 <= nm <= 2 * world + 2
            moved = 0
            for k in range(nm):
                assert ln[k] > 0 and src[k] >= total and dst[k] + ln[k] <= total       # from beyond the total into a gap below it
                lst[dst[k]:dst[k] + ln[k]] = lst[src[k]:src[k] + ln[k]]
                moved += int(ln[k])
            assert (np.sort(lst[:total]) == want).all(), (world, trial)
            assert moved <= int(begin[-1]) - total                                      # never more than the slack
    assert lib.ibvh_peer_compact_plan(0, None, None, None, None, None, 0) == -1
