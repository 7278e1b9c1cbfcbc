for v in "" $(ls implicitbvh.jl_b200/lib/variants/*.so 2>/dev/null); do
  IBVH_B200_LIB=${v:+$PWD/$v} timeout 200 python bench.py --steps 20 --warmup 5 --no-rays --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/v.json
  python tools/scale_line.py gpurun_out/v.json "lib=${v:-default}"
done
