set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_heads.py -x -q 2>&1 | tail -15
timeout 300 python tools/heads_bench.py 2>&1 | tail -3 | tee gpurun_out/heads_bench.json
