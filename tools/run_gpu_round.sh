set -x
timeout 600 python -m pytest tests/test_gpu_opacity.py tests/test_gpu_heads.py -x -q 2>&1 | tail -3
timeout 300 python tools/heads_bench.py 2>&1 | tail -1 | tee gpurun_out/heads_bench.json
