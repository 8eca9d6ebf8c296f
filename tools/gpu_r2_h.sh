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
for o in "--short-rows 4" "--short-rows 4 --opt item_nnz=128" "--short-rows 4 --opt item_nnz=512"; do
  echo "== arxiv $o"; $B --shape arxiv --steps 50 $o 2>>gpurun_out/h_err.log | show
  echo "== products $o"; $B --shape products $o 2>>gpurun_out/h_err.log | show
done
for lib in pygim_b200/libbackend_pim_hc768.so pygim_b200/libbackend_pim_hc1024.so; do
  echo "== $lib clustered tiles"; PYGIM_LIB_PATH=$lib $B --clustered --reorder tiles --tile-super-nnz 131072 --hot-k 1024 2>>gpurun_out/h_err.log | show
done
tail -3 gpurun_out/h_err.log
