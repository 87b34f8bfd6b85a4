set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/r1_bench_ours.json
for n in 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/r1_bench_ours_${n}gpu.json
done
for f in r1_bench_ours r1_bench_ours_2gpu r1_bench_ours_4gpu r1_bench_ours_8gpu; do python - <<PY
import json
d=json.load(open('gpurun_out/$f.json')); print('$f', round(d['value'],1), round(d['ms_per_step'],4), round(d['ms_per_step_median_rank0'],4), round(d['ms_per_step_max_rank0'],3), round(d['e2e']['value'],1), (d.get('cpu_baseline') or {}).get('value'))
PY
done
