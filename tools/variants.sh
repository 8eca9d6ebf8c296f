#!/usr/bin/env bash
# Tuning helper (GPU box): time the Reddit sweep with every libbackend_pim*.so variant present.
#   tools/variants.sh [extra bench.py flags]
cd "$(dirname "$0")/.."
for lib in pygim_b200/libbackend_pim*.so; do
  echo "== $lib $*"
  PYGIM_LIB_PATH=$lib python bench.py --steps 5 --warmup 3 --no-cpu --no-check --no-e2e "$@" 2>&1 | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.0f GFLOP/s' % d['value'], d['plan']['ds_parts']); [print('  H=%3d %.3f ms  %.0f GFLOP/s  gather %.1f TB/s' % (p['hidden'], p['kernel_ms'], p['gflops'], p['gather_gbs']/1e3)) for p in d['per_hidden']]"
done
