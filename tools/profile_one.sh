ncu --set full --clock-control none --import-source on -k regex:"visible_sort" -s 3 -c 1 -f -o gpurun_out/r2d_vs python bench.py --profile-only --steps 1 --warmup 3 > gpurun_out/r2d_ncu.log 2>&1
