#!/bin/bash
# usage: scratch/build_variant.sh NAME [-DPART_NT=128 ...]   -> scratch/variants/lib_NAME.so (config-2 kernels only, ~20 s)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p scratch/variants
python - "$name" "$@" <<'PY'
import sys, rimu_b200
name, defs = sys.argv[1], sys.argv[2:]
print("built", rimu_b200.build(defines=["-DRIMU_TUNE_ONLY_MOM1D"] + defs, out=f"scratch/variants/lib_{name}.so", kinds=[9]))
PY
