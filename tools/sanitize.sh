#!/bin/bash
# compute-sanitizer over the small-shape GPU tests (run on the GPU box):  bash tools/sanitize.sh [outdir]
# memcheck: every kernel of the path (update TMA ring, plan cluster / DSMEM, one-kernel resample, pick, draw, utility,
# multinomial parity path, batched engine, sweeper); racecheck / synccheck: the kernels with hand-rolled shared-memory
# protocols (mbarrier ring, per-warp marks, named barriers, cluster exchange); the device-side resample test (gated
# launches chained by programmatic dependent launch) is part of the selection.
out=${1:-gpurun_out}
mkdir -p "$out"
SEL_SMALL='tests/test_gpu_early_select.py::test_picked_draws_are_the_offspring tests/test_gpu_early_select.py::test_cycle_entry_equals_the_stepwise_path tests/test_gpu_early_select.py::test_shard_whose_first_slot_is_not_a_multiple_of_4 tests/test_gpu_device_test.py::test_device_test_equals_the_host_decision tests/test_gpu_device_test.py::test_device_test_not_used_where_it_does_not_apply tests/test_gpu_device_test.py::test_result_delivery_variants_agree tests/test_gpu_parity.py::test_golden_trajectory tests/test_gpu_batched.py::test_batched_lorentzian_matches_single_engines tests/test_gpu_sweeper.py::test_multi_point_update_matches_point_by_point'
K='not 3_000_001 and not 250_007 and not 1000003 and not 50000'
for tool in memcheck racecheck synccheck; do
  echo "== $tool" > "$out/sanitizer_$tool.log"
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --target-processes all --print-limit 20 \
      python -m pytest $SEL_SMALL -m gpu -x -q -k "$K" -p no:cacheprovider >> "$out/sanitizer_$tool.log" 2>&1
  echo "exit code $?" >> "$out/sanitizer_$tool.log"
  grep -E "ERROR SUMMARY|passed|failed|exit code" "$out/sanitizer_$tool.log" | tail -4
done
