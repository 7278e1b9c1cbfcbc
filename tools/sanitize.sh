#!/bin/bash
# compute-sanitizer over the small GPU parity tests (every kernel family runs at these sizes: the pyramid traversal starts
# from an all-pairs top level on small trees). memcheck on the wider subset, racecheck + synccheck on a narrow one.
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh r2w'
out=gpurun_out/${1:-sanitize}
mkdir -p $out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
SMALL="five_spheres or shuffled_doctest or pair_and_ray or unordered_structure or morton_bit_exact or morton_quirks or build_bit_exact or build_box or many_ties or built_level or stage_entry or single_all_start or counts_cache2 or other_index or single_box or degenerate or shards_concatenate or pair_all_start or self_equivalence or rays_small or float64 or mixed_float64 or deferred or reference_shaped or sidecar or triangles or bfs_doctests or bfs_rays"
[ "${SKIP_MEMCHECK:-0}" = 1 ] || timeout ${MEMCHECK_TIMEOUT:-900} compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 30 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SMALL" > $out/memcheck.txt 2>&1
echo "memcheck rc=$?" | tee -a $out/summary.txt
NARROW="${NARROW:-five_spheres or pair_and_ray or many_ties or single_box or rays_small or sidecar or bfs_doctests}"
for tool in racecheck synccheck; do
    timeout ${RACE_TIMEOUT:-600} compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 30 \
        python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$NARROW" > $out/$tool.txt 2>&1
    echo "$tool rc=$?" | tee -a $out/summary.txt
done
grep -h -E "ERROR SUMMARY|passed|failed|Invalid|Race reported|hazard" $out/*.txt | sort | uniq -c | head -40
