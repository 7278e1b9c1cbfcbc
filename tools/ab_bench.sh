#!/bin/bash
# A/B of library variants on the headline step: prints per-kernel ms for each (lib, env) combination
out=$1; shift
mkdir -p $out
run() { # name lib env...
  name=$1; lib=$2; shift 2
  env IBVH_B200_LIB=$lib "$@" python bench.py --steps 10 --warmup 3 --no-rays --no-cpu-baseline --workloads none > $out/$name.json 2> $out/$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$out/$name.json") if l.startswith("{")][-1])
    pk=d["roofline"]["per_kernel_ms"]
    print("$name", "step %.3f ms" % d["ms_per_step"], "build %.3f trav %.3f" % (d["roofline"]["build_ms_per_step"], d["roofline"]["traversal_ms_per_step"]), {k:v for k,v in pk.items() if v>0.03})
except Exception as e:
    print("$name FAILED", e, open("$out/$name.err").read()[-400:])
PY
}
L=$PWD/implicitbvh.jl_b200/lib
run tma_all $L/libibvh_b200.so
run tma_tile_only $L/libibvh_b200.so IBVH_PYR_TMA=0
run tma_refine_only $L/variants/libibvh_tma0.so
run tma_none $L/variants/libibvh_tma0.so IBVH_PYR_TMA=0
run tma_all_minb6 $L/variants/libibvh_tma_minb6.so
