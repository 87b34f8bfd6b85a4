set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graphs.py -x -q 2>&1 | tail -12
timeout 600 python tools/graph_bench.py 2>&1 | tail -2 | tee gpurun_out/graph_bench.json
