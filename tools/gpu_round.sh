#!/bin/bash
# One gpurun call: GPU parity tests, bench line, per-kernel timings, ncu launch list and one
# --set full capture of the hot kernels.  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r1_04'
# Env: SKIP_TESTS=1 / SKIP_NCU=1 / SKIP_FULL=1 to shorten the call.
tag=${1:-run}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
nproc >> $out/${tag}_gpu.txt

if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
  echo "pytest exit $?" >> $out/${tag}_pytest.log
  tail -5 $out/${tag}_pytest.log
fi

timeout 600 python bench.py --steps 30 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench exit $?"; cat $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
timeout 300 python tools/bench_kernels.py --iters 10 > $out/${tag}_kernels.json 2> $out/${tag}_kernels.err
echo "bench_kernels exit $?"; tail -3 $out/${tag}_kernels.err

if [ -z "$SKIP_NCU" ]; then
  # launch list of the bench command (eager launches so every kernel is visible)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
      python bench.py --steps 2 --warmup 3 --graph off --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
  echo "ncu launches exit $?"
  python tools/summarize_ncu.py launches $out/${tag}_launches.csv > $out/${tag}_launches.md 2>&1
  head -20 $out/${tag}_launches.md
fi
if [ -z "$SKIP_FULL" ]; then
  # one full capture of our kernels in one steady-state training step (skip the first 4 steps = 40 launches)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_grid|k_tc|k_march|k_composite' --launch-skip 36 -c 9 \
      -f -o $out/${tag}_full python bench.py --steps 2 --warmup 3 --graph off --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1
  echo "ncu full exit $?"
  python tools/summarize_ncu.py full $out/${tag}_full.ncu-rep > $out/${tag}_full.md 2>&1
  cat $out/${tag}_full.md
fi
