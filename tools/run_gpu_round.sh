set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/dropin_bench.py 2>&1 | tail -1
OCRF_SPECULATIVE=0 timeout 300 python tools/dropin_bench.py 2>&1 | tail -1
OCRF_BENCH_EXACT=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('exact', d['value'], d['ms_per_step'], d['e2e']['value'])"
