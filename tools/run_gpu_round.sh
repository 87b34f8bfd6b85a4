set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/r1_bench_ours.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r1_bench_reference.json
for f in r1_bench_ours r1_bench_reference; do python - <<PY
import json
d=json.load(open('gpurun_out/$f.json')); print('$f', d['value'], d['ms_per_step'], d.get('ms_per_step_median_rank0'), d['e2e']['value'], (d.get('cpu_baseline') or {}).get('value'), d.get('clocks'))
PY
done
