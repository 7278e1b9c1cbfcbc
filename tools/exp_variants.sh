for v in "" implicitbvh.jl_b200/lib/variants/libibvh_fan2.so; do
  IBVH_B200_LIB=${v:+$PWD/$v} timeout 200 python bench.py --steps 20 --warmup 5 --no-rays --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/v.json
  python tools/scale_line.py gpurun_out/v.json "lib=${v:-default}"
  python -c "
import json; d=json.load(open('gpurun_out/v.json')); print(d['roofline']['per_kernel_ms'], d['roofline'].get('launches_per_step'))"
done
