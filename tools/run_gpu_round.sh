set -x
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'])"; done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
