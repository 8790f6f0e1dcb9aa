#!/bin/bash
# usage: scratch/sweep_env.sh VAR v1 v2 ...   -- bench.py once per value of an environment knob; prints ms/step and phases
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-replicas 1 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$var=$v', round(d['ms_per_step'],4), d['extra']['phase_ms_per_step'])"
done
