set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_voxel_color.py -x -q 2>&1 | tail -12
timeout 900 python tools/voxel_color_bench.py 2>&1 | tail -1 | tee gpurun_out/voxel_color_bench.json
