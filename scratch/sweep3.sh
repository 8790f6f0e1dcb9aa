#!/bin/bash
cd "$(dirname "$0")/.."
for rep in 1 2; do
for name in "$@"; do
  RIMU_BENCH_SKIP_PREFLIGHT=1 RIMU_B200_LIB=$PWD/scratch/variants/lib_$name.so timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-replicas 1 --long-steps 200 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],4), d['extra']['phase_ms_per_step'], d['extra']['long_run'])"
done
done
