#!/bin/bash
# Final profiling call of round 2: launch list of the bench step (same command as the bench line, eager launches so that every kernel is
# visible) and one `ncu --set full` capture of a steady-state training step.   gpurun --timeout 1500 -- 'bash tools/gpu_profile_final.sh r2_56'
tag=${1:-r2_final}
out=gpurun_out
mkdir -p $out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --graph off --no-cpu-baseline --skip strong_scaling,event_step,run_variant,gpu_bar,render,extra_state > $out/${tag}_ncu_bench.log 2>&1
echo "ncu launches exit $?"
python tools/summarize_ncu.py launches $out/${tag}_launches.csv > $out/${tag}_launches.md 2>&1
head -24 $out/${tag}_launches.md
timeout 700 ncu --set full --clock-control none --import-source on -k regex:'k_grid|k_tc|k_march|k_composite|k_adam|k_finish|k_near|k_occ' --launch-skip 160 -c 20 \
    -f -o $out/${tag}_full_step python bench.py --steps 2 --warmup 3 --graph off --no-cpu-baseline --skip strong_scaling,event_step,run_variant,gpu_bar,render,extra_state > $out/${tag}_ncu_full.log 2>&1
echo "ncu full step exit $?"
python tools/summarize_ncu.py full $out/${tag}_full_step.ncu-rep > $out/${tag}_full_step.md 2>&1
head -30 $out/${tag}_full_step.md
