set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/r1_bench_ours_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r1_bench_ours_2gpu.json')); print(d['value'], d['ms_per_step'], d['ms_per_step_median_rank0'], d['e2e'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['ms_per_step_median_rank0'], d['e2e']['value'])"
