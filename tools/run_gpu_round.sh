set -x
mkdir -p gpurun_out
for pdl in 0 1; do
OCRF_PDL=$pdl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$pdl bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PDL=$pdl', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
