#!/bin/bash
# compute-sanitizer over the small-shape GPU tests (run on the GPU box):  bash tools/sanitize.sh [outdir]
# memcheck: every kernel of the path (update TMA ring, plan cluster / DSMEM, one-kernel resample, pick, draw, utility,
# multinomial parity path, batched engine, sweeper); racecheck / synccheck: the kernels with hand-rolled shared-memory
# protocols (mbarrier ring, per-warp marks, named barriers, cluster exchange).
out=${1:-gpurun_out}
mkdir -p "$out"
SEL_SMALL='tests/test_gpu_early_select.py::test_picked_draws_are_the_offspring tests/test_gpu_early_select.py::test_cycle_entry_equals_the_stepwise_path tests/test_gpu_early_select.py::test_shard_with_an_odd_first_slot_stores_the_same_cloud tests/test_gpu_parity.py::test_golden_trajectory tests/test_gpu_batched.py::test_batched_lorentzian_matches_single_engines tests/test_gpu_sweeper.py::test_multi_point_update_matches_point_by_point'
K='not 3_000_001 and not 250_007'
for tool in memcheck racecheck synccheck; do
  echo "== $tool" > "$out/sanitizer_$tool.log"
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --target-processes all --print-limit 20 \
      python -m pytest $SEL_SMALL -m gpu -x -q -k "$K" -p no:cacheprovider >> "$out/sanitizer_$tool.log" 2>&1
  echo "exit code $?" >> "$out/sanitizer_$tool.log"
  grep -E "ERROR SUMMARY|passed|failed|exit code" "$out/sanitizer_$tool.log" | tail -4
done
