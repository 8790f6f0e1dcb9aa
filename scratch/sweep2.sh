#!/bin/bash
# usage: scratch/sweep2.sh -- config 2 (bench.py) for the MOM1D-only variants, config 4 (bench_configs, 2e7 walkers) for the full-build variants,
# and the parity tests on the full build with async parents
cd "$(dirname "$0")/.."
for rep in 1 2; do
for name in sync async async2; do
  RIMU_B200_LIB=$PWD/scratch/variants/lib_$name.so timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-replicas 1 --long-steps 200 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],4), d['extra']['phase_ms_per_step'], d['extra']['long_run'])"
done
done
for name in fullsync fullasync w2cap1024; do
  RIMU_B200_LIB=$PWD/scratch/variants/lib_$name.so timeout 200 python bench_configs.py --configs 4,7 --walkers 2e7 --steps 20 2>/dev/null | grep "^{" | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$name', d['config'][:24], round(d['ms_per_step'],4), d['phase_ms'], d['buckets_per_gpu'], d['max_bucket_fill'])"
done
RIMU_B200_LIB=$PWD/scratch/variants/lib_fullasync.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_initiators.py tests/test_gpu_advance.py -x -q 2>&1 | tail -5
