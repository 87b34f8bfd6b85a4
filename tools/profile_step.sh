#!/bin/bash
# One full-counter ncu capture of every ocrf kernel of ONE bench step (the step after 3 warm-up steps: 9 kernels per
# step), with source correlation, plus the launch list of a short bench run.  Outputs under gpurun_out/ (tag = $1).
tag=${1:-r2}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"ocrf|visible_sort|ms_|render_|preprocess_|clear_" -s 27 -c 9 \
    -f -o gpurun_out/${tag}_step python bench.py --profile-only --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --profile-only --steps 2 --warmup 3 > gpurun_out/${tag}_launches.log 2>&1
ls -la gpurun_out/${tag}_step.ncu-rep
