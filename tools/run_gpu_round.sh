set -x
timeout 900 python tools/config_sanity.py 2>&1 | tail -4 | cut -c1-260
