#!/bin/bash
# First GPU call of the next round: run what was written after round 1's GPU budget was spent.
#   gpurun --timeout 600 -- 'bash tools/gpu_next_round_first.sh r2_01'
tag=${1:-r2_01}
out=gpurun_out
mkdir -p $out
# 1. paired-load gather (forward mode 3): bit-identity vs the generic kernel and the reference build, then its time next to mode 1
ENERF_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_encoders.py -q -k hoisted > $out/${tag}_pytest_experimental.log 2>&1
echo "experimental tests exit $?"; tail -3 $out/${tag}_pytest_experimental.log
timeout 200 python tools/bench_kernels.py --only grid --iters 10 > $out/${tag}_kernels_grid.json 2> $out/${tag}_kernels_grid.err
echo "bench_kernels exit $?"; python - <<PY
import json
r = json.load(open("$out/${tag}_kernels_grid.json"))
for k in ("grid_fwd_float16", "grid_fwd_float16_paired_loads"):
    print(k, r.get(k))
PY
# 2. the whole parity suite and the contract line on this round's starting code
timeout 600 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 $out/${tag}_pytest.log
timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench exit $?"; cut -c1-300 $out/${tag}_bench.json
