#!/bin/bash
for cfg in "1 0 12" "0 0 12" "1 0 8" "1 0 16" "0 0 16" "1 1 12"; do
  set -- $cfg
  echo "== tile=$1 sort=$2 bpsm=$3"
  NCB_RAY_TILE=$1 NCB_RAY_SORT=$2 NCB_RAY_BPSM=$3 python bench.py --rays-only --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('Mrays/s',round(d['value'],1),'ms',round(d['ms_per_batch'],4),'e2e',round(d['e2e']['value'],1))"
done
