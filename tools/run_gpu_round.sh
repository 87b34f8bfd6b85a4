set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/r1_bench_ours.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/r1_bench_ours_2gpu.json
for f in r1_bench_ours r1_bench_ours_2gpu; do python - <<PY
import json
d=json.load(open('gpurun_out/$f.json')); print('$f', d['value'], d['ms_per_step'], d['ms_per_step_median_rank0'], d['ms_steps_rank0'][:3], d['e2e']['value'], d.get('cpu_baseline',{}).get('value'))
PY
done
