#!/bin/bash
# A/B of the blend-kernel variants on the bench workload (stage timers of bench.py): prints one line per variant
run() {
  env "$@" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$*', 'step %.4f ms' % d['ms_per_step'], {k: v['ms'] for k, v in d['stages'].items()})"
}
run OCRF_FWD_PACKED=0 OCRF_BWD_PACKED=0
run OCRF_FWD_PACKED=1 OCRF_BWD_PACKED=0
run OCRF_FWD_PACKED=1 OCRF_FWD_PPT=4 OCRF_BWD_PACKED=0
run OCRF_FWD_PACKED=1 OCRF_BWD_PACKED=1
run OCRF_FWD_PACKED=1 OCRF_BWD_PACKED=1 OCRF_BWD_PPT=2
