set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_heads.py -x -q 2>&1 | tail -4
timeout 600 python tools/heads_bench.py 2>&1 | tail -1 | tee gpurun_out/heads_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"gaussian_heads" -s 30 -c 6 --csv --log-file gpurun_out/heads_launches.csv python tools/heads_bench.py > /dev/null 2>&1
grep -v "^==" gpurun_out/heads_launches.csv | cut -d, -f5,13,15 | tail -24
