#!/bin/bash
cd "$(dirname "$0")/.."
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2_pytest_gpu_final.log
timeout 300 python bench.py 2>/dev/null | grep "^{" | tee gpurun_out/r2_bench1_final2.log | cut -c1-200
ncu --profile-from-start off --set full --clock-control none --import-source on -c 3 -f -o gpurun_out/r2_final2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-replicas 1 --long-steps 0 > gpurun_out/r2_ncu_final2.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_final2_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-replicas 1 --long-steps 0 > gpurun_out/r2_ncu_launches2.log 2>&1
for name in sync async async2; do
  RIMU_BENCH_SKIP_PREFLIGHT=1 RIMU_B200_LIB=$PWD/scratch/variants/lib_$name.so timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-replicas 1 --long-steps 200 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],4), d['extra']['phase_ms_per_step'], d['extra']['long_run'])"
done 2>&1 | tee gpurun_out/r2_sweep_async_w1.log
