#!/bin/bash
# usage: scratch/sweep.sh name1 name2 ...  -> one compact line per variant (bench.py config 2)
cd "$(dirname "$0")/.."
for v in "$@"; do
  RIMU_B200_LIB=$PWD/scratch/variants/lib_$v.so timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -n 1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', 'ms/step %.4f'%d['ms_per_step'], d['extra']['phase_ms_per_step'], 'e2e %.3g'%d['e2e']['value'])" 2>&1 | tail -n 1
done
