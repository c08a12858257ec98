#!/bin/bash
# compute-sanitizer over the kernels touched last in round 2 (the fused inference field incl. its device-side row count, the inference
# marcher that writes every row of its slots, the float rows of the occupancy bounds).   gpurun --timeout 110 -- 'bash tools/sanitize_r2_67.sh r2_67'
tag=${1:-san}
out=gpurun_out
mkdir -p $out
SEL='test_device_side_row_count_limits_the_rows_evaluated and ([0] or [37]) or test_fused_field_is_the_unfused_chain_bit_for_bit and ([127] or [5000]) or test_inference_march_writes_every_row_of_its_slots or test_bounded_march_emits_the_samples_of_the_exhaustive_march and corner'
export ENERF_SANITIZER=1
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do
  extra=""
  [ "$tool" = "memcheck" ] && extra="--leak-check no --report-api-errors no"
  timeout ${SAN_TIMEOUT:-45} compute-sanitizer --tool $tool $extra --print-limit 30 --error-exitcode 99 \
      python -m pytest tests/test_gpu_field_infer.py tests/test_gpu_raymarching.py -m gpu -q -x -k "$SEL" -p no:cacheprovider > $out/${tag}_sanitizer_${tool}.log 2>&1
  echo "$tool exit $?"
  grep -E "ERROR SUMMARY|passed|failed|error" $out/${tag}_sanitizer_${tool}.log | tail -4
done
