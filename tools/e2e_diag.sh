#!/bin/bash
# where does the N=2 end-to-end step go? (diagnosis; IBVH_E2E_SKIP drops one leg of the pipelined loop)
out=$1; mkdir -p $out
for skip in none h2d allgather d2h "h2d,d2h" "h2d,allgather,d2h"; do
  IBVH_E2E_SKIP=$skip python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29580 bench.py --gpus 2 --steps 10 --warmup 3 --workloads none --no-rays > $out/e2e_$skip.json 2> $out/e2e_$skip.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$out/e2e_$skip.json") if l.startswith("{")][-1])
    print("skip=$skip", "step %.3f" % d["ms_per_step"], "e2e pipelined %.2f ms sequential %.2f ms" % (d["e2e"]["ms_per_step"], d["e2e"]["sequential_ms_per_step"]))
except Exception as e:
    print("skip=$skip failed", e)
PY
done
