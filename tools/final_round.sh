#!/bin/bash
# End-of-round evidence run on one B200:   gpurun --timeout 1500 -- 'bash tools/final_round.sh r2ah'
#   pytest -m gpu, the contract bench line, the ncu launch list of a short bench run, one ncu --set full capture of the
#   (unordered) leaf-tile kernel. Timings in the bench line are CUDA events; nothing printed under ncu is a bench value.
out=gpurun_out/${1:-final}
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest.txt 2>&1; tail -2 $out/pytest.txt
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; tail -2 $out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-rays --no-cpu-baseline --workloads none > $out/b_ncu.log 2>&1
python tools/launch_list_summary.py $out/launches_raw.csv $out/launch_list_one_step.csv; cat $out/launch_list_one_step.csv
ncu --set full --clock-control none -k regex:pyr_leaf_tile_kernel -c 2 -o /tmp/tile python tools/ncu_unordered.py > $out/ncu_tile.log 2>&1
python tools/ncu_summary.py /tmp/tile.ncu-rep $out/ncu_tile_f32x2.csv; cut -c1-400 $out/ncu_tile_f32x2.csv
