#!/bin/bash
# timing experiments for the SimRank kernel (GW_SR_DEBUG: 1 = no inserts, 2 = no tier 2, 4 = no top-k)
for d in 0 1 2 4 6; do
  echo "debug=$d"
  GW_SR_DEBUG=$d python bench.py --workload simrank --steps 2 --warmup 1 --ba-nodes 1000000 --queries-per-step 2048 --no-cpu-baseline --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   ', round(d['value']),'q/s', round(d['ms_per_step'],2),'ms')"
done
