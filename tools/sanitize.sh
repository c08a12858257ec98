#!/bin/bash
# compute-sanitizer over reduced-size GPU tests (race detection / sanitizers).  One gpurun call:
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh r2_14'
# memcheck (out-of-bounds / misaligned accesses, leaks of device-side errors), racecheck (shared-memory hazards) and synccheck
# (invalid barrier usage) each run the same selection; logs land in gpurun_out/<tag>_sanitizer_<tool>.log.
tag=${1:-san}
out=gpurun_out
mkdir -p $out
SEL='test_forward_inference_backward_vs_oracle and (1024 or 256-16) or test_density_head_matches_linear_chain and 128] or test_masked_color_matches_module_chain and 1-0.3 or test_grid_forward_hoisted_kernel_is_bit_identical_to_generic or test_near_far_bit_exact or test_march_rays_train_mean_count_overflow_drops_rays or test_inference_loop_primitives_match_oracle or test_density_grid_maintenance or test_composite_train or test_ordered_compaction_matches_nonzero and (4097 or 16]) or test_weighted_sum_and_row_moves or test_fused_field_matches_module_chain or test_sh_ or test_bounded_march_emits_the_samples_of_the_exhaustive_march and (corner or sparse) or test_finish_rays_is_the_aten_expression and per_ray or test_grid_module_input_mapping_inside_the_kernels_is_bit_identical and 0.75'
export ENERF_SANITIZER=1
for tool in memcheck racecheck synccheck; do
  extra=""
  [ "$tool" = "memcheck" ] && extra="--leak-check no --report-api-errors no"
  timeout 1200 compute-sanitizer --tool $tool $extra --print-limit 30 --error-exitcode 99 \
      python -m pytest tests/test_gpu_ffmlp.py tests/test_gpu_torch_field.py tests/test_gpu_encoders.py tests/test_gpu_raymarching.py tests/test_gpu_occupancy.py tests/test_gpu_renderer.py \
      -m gpu -q -x -k "$SEL" -p no:cacheprovider > $out/${tag}_sanitizer_${tool}.log 2>&1
  echo "$tool exit $?"
  grep -E "ERROR SUMMARY|passed|failed|error" $out/${tag}_sanitizer_${tool}.log | tail -4
done
