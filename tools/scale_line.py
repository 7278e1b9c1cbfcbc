import json, sys
d = json.load(open(sys.argv[1]))
pk = d["roofline"]["per_kernel_ms"]
print(sys.argv[2] if len(sys.argv) > 2 else "", d["n_gpus"], "value", round(d["value"] / 1e6, 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"] / 1e6, 1),
      {k: pk[k] for k in pk if k.startswith(("pyr_leaf", "pyr_refine", "peer"))})
