#!/bin/bash
for b in 5 6 8; do for pq in "0.25 4" "1 1"; do set -- $pq
  echo "MINB=$b p=$1 q=$2"
  GW_CN_MINB=$b python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --p $1 --q $2 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   ', round(d['value']/1e9,2),'G steps/s', round(d['ms_per_step'],3),'ms')"
done; done
