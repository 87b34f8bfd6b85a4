set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -x -q 2>&1 | tail -4
timeout 600 python tools/config_sanity.py 2>&1 | grep "config4" | cut -c1-260
