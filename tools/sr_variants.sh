#!/bin/bash
# A/B timing of the SimRank kernel variants built by tools/build_sr_variants.sh
for lib in "" $(ls tools/variants/libgw_*.so); do
  echo "lib=${lib:-default}"
  GW_LIB_OVERRIDE=$lib python bench.py --workload simrank --steps 3 --warmup 1 --ba-nodes ${BA:-1000000} --queries-per-step 2048 --no-cpu-baseline --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   ', round(d['value']),'q/s', round(d['ms_per_step'],2),'ms', 'slow', d.get('slow_path_queries_last_step'))"
done
