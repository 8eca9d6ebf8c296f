#!/bin/bash
# round-2 GPU call B: whole parity suite + kernel variants on uniform / products / clustered graphs
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
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --ignore=tests/test_gpu_fullsize.py > gpurun_out/b_tests.log 2>&1
echo "exit $?" >> gpurun_out/b_tests.log
tail -15 gpurun_out/b_tests.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-clustered --no-products"
for lib in pygim_b200/libbackend_pim.so pygim_b200/libbackend_pim_nv0.so pygim_b200/libbackend_pim_nv1.so; do
  for shape in reddit products; do
    echo "== $lib $shape"
    PYGIM_LIB_PATH=$lib $B --shape $shape 2>>gpurun_out/b_err.log | tee -a gpurun_out/b_variants.jsonl | show
  done
done
echo "== clustered, natural order"
$B --clustered 2>>gpurun_out/b_err.log | tee -a gpurun_out/b_clustered.jsonl | show
for opts in "" "--opt cta_threads=1024" "--opt max_g=8" "--opt cta_threads=1024 --opt max_g=8" "--opt super_nnz=16384" "--opt super_nnz=16384 --opt max_g=8"; do
  echo "== clustered + reorder cluster $opts"
  $B --clustered --reorder cluster $opts 2>>gpurun_out/b_err.log | tee -a gpurun_out/b_clustered.jsonl | show
  echo "== same, shuffle index delivery"
  PYGIM_LIB_PATH=pygim_b200/libbackend_pim_nv0.so $B --clustered --reorder cluster $opts 2>>gpurun_out/b_err.log | tee -a gpurun_out/b_clustered.jsonl | show
done
tail -5 gpurun_out/b_err.log
