#!/bin/bash
# Last call of the round (short budget): full GPU parity suite, the two-stream overlap probe, one bench line with the
# pipelined backward at the probe's best setting.  Everything lands in gpurun_out/<tag>_*.
tag=${1:-final}
out=gpurun_out
mkdir -p $out
timeout 150 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -15 $out/${tag}_pytest.log
timeout 70 python tools/overlap_probe.py --iters 8 --out $out/${tag}_overlap_probe.json > $out/${tag}_overlap_probe.log 2>&1
echo "probe exit $?"; tail -3 $out/${tag}_overlap_probe.log | cut -c1-1500
best=$(python - <<PY
import json
try:
    r = json.load(open("$out/${tag}_overlap_probe.json"))
    ok = [c for c in r["bwd"] if "ms" in c and max(c["max_err_rel_to_max"].values()) < 1e-3]
    b = min(ok, key=lambda c: c["ms"])
    print(b["chunks"], b["mlp_ctas"], b["scatter_block"])
except Exception:
    print(4, 120, 256)
PY
)
set -- $best
echo "bench with pipeline chunks=$1 mlp_ctas=$2 scatter_block=$3"
ENERF_PIPELINE=1 ENERF_PIPELINE_CHUNKS=$1 ENERF_PIPELINE_MLP_CTAS=$2 ENERF_PIPELINE_SCATTER_BLOCK=$3 \
  timeout 70 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-render > $out/${tag}_bench_pipeline.json 2> $out/${tag}_bench_pipeline.err
echo "bench exit $?"; cut -c1-600 $out/${tag}_bench_pipeline.json; tail -3 $out/${tag}_bench_pipeline.err
