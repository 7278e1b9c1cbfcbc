#!/bin/bash
# ncu --set full of the contact traversal's tile and refine kernels (one launch each, after warm-up) with per-instruction
# (SASS) counters exported as CSV on the box: the .ncu-rep itself is too large to travel back.
#   gpurun -- 'bash tools/ncu_tile_source.sh r2t [kernel regex] [launches]'
out=gpurun_out/${1:-ncu_src}
mkdir -p $out
ncu --set full --profile-from-start off --clock-control none --import-source on -k regex:"${2:-pyr_leaf_tile_kernel}" -c ${3:-2} \
    -o /tmp/tile python tools/ncu_step.py > $out/run.log 2>&1
ncu -i /tmp/tile.ncu-rep --page raw --csv > $out/raw.csv 2>/dev/null
ncu -i /tmp/tile.ncu-rep --page source --csv --print-source sass > $out/source_sass.csv 2>$out/source.err || \
ncu -i /tmp/tile.ncu-rep --page source --csv > $out/source_sass.csv 2>>$out/source.err
ls -la $out /tmp/tile.ncu-rep
