set -x
for i in 1 2 3; do timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['ms_per_step_median_rank0'], d['ms_per_step_max_rank0'], d['ms_steps_rank0'][:4], d['e2e']['value'], d['clocks'])"; done
