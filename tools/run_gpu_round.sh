set -x
timeout 300 python tools/dropin_bench.py 2>&1 | tail -1 | tee gpurun_out/r1_dropin_bench.json
