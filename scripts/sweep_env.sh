#!/bin/bash
# bash scripts/sweep_env.sh VAR "v1 v2 v3" -> bench stage times for each value
VAR=$1; VALS=$2
for v in $VALS; do
  echo "== $VAR=$v"
  env $VAR=$v python bench.py --steps 5 --warmup 3 --no-cpu --no-rays 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms_per_step',round(d['ms_per_step'],3),' '.join(f\"{s['stage']}={s['ms']}\" for s in d['stages_ms']))"
done
