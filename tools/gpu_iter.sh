#!/bin/bash
# Short GPU iteration: targeted tests + per-kernel timings + one bench line.  gpurun -- 'bash tools/gpu_iter.sh tag'
tag=${1:-it}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q -k "${TESTS:-ffmlp or field or grid or golden or renderer}" > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -15 $out/${tag}_pytest.log
timeout 300 python tools/bench_kernels.py --iters 10 --only ${ONLY:-mlp,grid} > $out/${tag}_kernels.json 2> $out/${tag}_kernels.err
echo "bench_kernels exit $?"; tail -3 $out/${tag}_kernels.err
python - <<PY
import json
d=json.load(open("$out/${tag}_kernels.json"))
for k,v in d.items():
    if isinstance(v,dict) and 'ms' in v: print(f"{k:45s} {v['ms']:.3f} ms  frac={v.get('frac')}")
PY
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --skip ${SKIP:-gpu_bar} > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench exit $?"; tail -3 $out/${tag}_bench.err
python - <<PY
import json
d=json.load(open("$out/${tag}_bench.json"))
print("ms_per_step", d["ms_per_step"], "rays/s", d["value"], "e2e", d["e2e"]["value"])
for n,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]['ms_per_step']): print(f"  {n:40s} {v['ms_per_step']:.3f} ms  frac={v.get('frac')}")
PY
