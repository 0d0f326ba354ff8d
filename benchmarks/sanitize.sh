#!/bin/bash
# compute-sanitizer over the GPU tests that exercise the look-back plans (plan_pairs_kernel, plan_lists_kernel), the
# peer exchange (one-rank self-exchange) and the tcgen05 / TMA kernels.  Run on the GPU box:
#   gpurun --timeout 1500 -- 'bash benchmarks/sanitize.sh'
# Logs land in gpurun_out/san/; the summaries kept for review are copied to profiles/ by hand.
set -u
OUT=gpurun_out/san
mkdir -p "$OUT"
SUBSET='test_plan_scans_across_many_ctas or test_sharded_step_with_a_one_rank_exchange or test_tensor_core_tiles_match_ffma_tiles_and_oracle or test_tensor_core_fused_argmax_matches_simt or (test_tensor_core_coarse_scores_have_fp32_accuracy and 768) or test_list_major_auto_mode_full_search or test_small_batch_path or test_one_pass_selection_and_its_fallbacks'
LIMIT=${SAN_LIMIT:-420}
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do
    t0=$(date +%s)
    timeout "$LIMIT" compute-sanitizer --tool "$tool" --log-file "$OUT/$tool.log" --print-limit 40 \
        python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "$SUBSET" > "$OUT/$tool.pytest.log" 2>&1
    rc=$?
    t1=$(date +%s)
    echo "== $tool rc=$rc $((t1 - t0)) s"
    tail -n 3 "$OUT/$tool.pytest.log"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$OUT/$tool.log" | tail -n 2
done
