set -x
mkdir -p gpurun_out/golden
timeout 900 python -m pytest tests/test_gpu_bev_pool.py -x -q 2>&1 | tail -4
timeout 600 python tools/bev_pool_bench.py 2>&1 | tail -1 | tee gpurun_out/bev_pool_bench.json
BEV_B=1 timeout 600 python tools/bev_pool_bench.py 2>&1 | tail -1
