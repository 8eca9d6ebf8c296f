#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print('   value %.0f GFLOP/s  parity %s' % (d['value'], d.get('parity_all_ranks')))
    for p in d['per_hidden']: print('     H=%3d %.3f ms  %.0f GFLOP/s  gather %.1f TB/s  frac %.3f' % (p['hidden'], p['kernel_ms'], p['gflops'], p['gather_gbs']/1e3, p['frac_hbm']))
"; }
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-clustered --no-products"
for lib in pygim_b200/libbackend_pim.so pygim_b200/libbackend_pim_t768nv2.so pygim_b200/libbackend_pim_t512nv3.so pygim_b200/libbackend_pim_t512nv4.so; do
  echo "== $lib reddit"
  PYGIM_LIB_PATH=$lib $B 2>>gpurun_out/d_err.log | tee -a gpurun_out/d_variants.jsonl | show
  echo "== $lib clustered+reorder"
  PYGIM_LIB_PATH=$lib $B --clustered --reorder cluster 2>>gpurun_out/d_err.log | tee -a gpurun_out/d_variants.jsonl | show
  echo "== $lib products short_rows=0"
  PYGIM_LIB_PATH=$lib $B --shape products --short-rows 0 2>>gpurun_out/d_err.log | tee -a gpurun_out/d_variants.jsonl | show
done
tail -3 gpurun_out/d_err.log
