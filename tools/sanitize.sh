#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over a small slice of the GPU parity tests
mkdir -p gpurun_out
SEL='test_golden_digests_through_c_abi or (test_digest_parity_small_batch and 0-uniform) or (test_slot_witness_parity and 8-0) or test_bitdecomp_witness_parity or test_msm_parity_small or test_msm_edge or test_msm_fixed_base or test_store_basic or (test_fold_helpers_parity and 0) or (test_ntt_parity and 10-0) or test_two_pipelined'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -q --timeout 1400 -x -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1) | $(tail -1 gpurun_out/sanitizer_$tool.log)"
done
