set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -x -q 2>&1 | tail -3
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['stages'].items()})"; done
OCRF_NVCC_EXTRA="-DOCRF_PRE_MINB=6" python -m ocrfdet_b200.build --force | tail -1
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['stages'].items()})"; done
OCRF_NVCC_EXTRA="-DOCRF_PRE_MINB=4" python -m ocrfdet_b200.build --force | tail -1
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['stages'].items()})"; done
