set -x
for i in 1 2; do timeout 600 python tools/c5_probe.py 2>&1 | tail -3 | cut -c1-230; done
