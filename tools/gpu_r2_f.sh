#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print('   value %.0f GFLOP/s  parity %s' % (d['value'], d.get('parity_all_ranks')), d.get('reorder_stats'))
    for p in d['per_hidden']: print('     H=%3d %.3f ms  %.0f GFLOP/s  gather %.1f TB/s  frac %.3f' % (p['hidden'], p['kernel_ms'], p['gflops'], p['gather_gbs']/1e3, p['frac_hbm']))
"; }
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-clustered --no-products"
echo "== clustered + tiles"; $B --clustered --reorder tiles 2>>gpurun_out/f_err.log | tee -a gpurun_out/f.jsonl | show
echo "== clustered + tiles super 131072 k1024"; $B --clustered --reorder tiles --tile-super-nnz 131072 --hot-k 1024 2>>gpurun_out/f_err.log | tee -a gpurun_out/f.jsonl | show
for o in "--short-rows 2" "--short-rows 2 --opt item_nnz=96" "--short-rows 2 --opt item_nnz=48" "--short-rows 3 --opt item_nnz=96" "--short-rows 1 --opt item_nnz=96"; do
  echo "== arxiv $o"; $B --shape arxiv --steps 50 $o 2>>gpurun_out/f_err.log | show
done
for o in "--short-rows 2" "--short-rows 3" "--short-rows 3 --opt item_nnz=128" "--short-rows 3 --opt cta_threads=512" "--short-rows 2 --opt item_nnz=128"; do
  echo "== products $o"; $B --shape products $o 2>>gpurun_out/f_err.log | show
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/f_ncu_arxiv.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-clustered --no-products --no-check --shape arxiv --short-rows 2 > /dev/null 2>>gpurun_out/f_err.log
tail -3 gpurun_out/f_err.log
