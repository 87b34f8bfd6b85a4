set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/heads_bench.py 2>&1 | tail -1 | tee gpurun_out/heads_bench.json
for d in 1 1; do timeout 300 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:v['ms'] for k,v in d['stages'].items()})"; done
