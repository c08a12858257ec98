#!/bin/bash
# Round-2 profiling call: launch list of the bench step and `ncu --set full` captures of (a) one steady-state headline training step,
# (b) one step of the shipped-config run() path, (c) the occupancy-grid refresh.   gpurun --timeout 2400 -- 'bash tools/gpu_profile_r2.sh r2_18'
tag=${1:-r2_prof}
out=gpurun_out
mkdir -p $out
# launch list (eager launches so every kernel is visible; 2 timed steps after 3 warm-ups)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --graph off --no-cpu-baseline --skip strong_scaling,event_step,run_variant,gpu_bar,render,extra_state > $out/${tag}_ncu_bench.log 2>&1
echo "ncu launches exit $?"
python tools/summarize_ncu.py launches $out/${tag}_launches.csv > $out/${tag}_launches.md 2>&1
head -30 $out/${tag}_launches.md
# (a) full capture of one headline step: skip the sizing + warm-up steps
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_grid|k_tc|k_march|k_composite|k_adam|k_occ|k_compact|k_packbits' --launch-skip 160 -c 24 \
    -f -o $out/${tag}_full_step python bench.py --steps 2 --warmup 3 --graph off --no-cpu-baseline --skip strong_scaling,event_step,run_variant,gpu_bar,render,extra_state > $out/${tag}_ncu_full.log 2>&1
echo "ncu full step exit $?"
python tools/summarize_ncu.py full $out/${tag}_full_step.ncu-rep > $out/${tag}_full_step.md 2>&1
head -40 $out/${tag}_full_step.md
# (b) the run() path
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_grid|k_tc|k_color|k_uniform|k_compact|k_weighted|k_gather|k_scatter' --launch-skip 60 -c 20 \
    -f -o $out/${tag}_full_run python bench.py --only run_variant --graph off --steps 3 > $out/${tag}_ncu_run.log 2>&1
echo "ncu full run exit $?"
python tools/summarize_ncu.py full $out/${tag}_full_run.ncu-rep > $out/${tag}_full_run.md 2>&1
head -40 $out/${tag}_full_run.md
