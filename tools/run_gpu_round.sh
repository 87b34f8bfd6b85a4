set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/r1_bench_ours_8gpu.json
python - <<PY
import json
d=json.load(open('gpurun_out/r1_bench_ours_8gpu.json')); print(8, d['value'], d['ms_per_step'], d['ms_per_step_median_rank0'], d['ms_steps_rank0'][:6], d['clocks'], d['e2e']['value'])
PY
