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
for lib in pygim_b200/libbackend_pim.so pygim_b200/libbackend_pim_nv1.so pygim_b200/libbackend_pim_t768.so pygim_b200/libbackend_pim_t768nv3.so; do
  echo "== $lib reddit"
  PYGIM_LIB_PATH=$lib $B 2>>gpurun_out/c_err.log | tee -a gpurun_out/c_variants.jsonl | show
done
echo "== default, item_nnz 512"; $B --opt item_nnz=512 2>>gpurun_out/c_err.log | show
echo "== default, cta 512"; $B --opt cta_threads=512 2>>gpurun_out/c_err.log | show
echo "== default, no l2 persist"; $B --no-l2-persist 2>>gpurun_out/c_err.log | show
for sr in 2 0 1; do
  echo "== products short_rows=$sr"
  $B --shape products --short-rows $sr 2>>gpurun_out/c_err.log | tee -a gpurun_out/c_variants.jsonl | show
done
echo "== products short_rows=0 item_nnz=128"; $B --shape products --short-rows 0 --opt item_nnz=128 2>>gpurun_out/c_err.log | show
echo "== products short_rows=2 item_nnz=128"; $B --shape products --short-rows 2 --opt item_nnz=128 2>>gpurun_out/c_err.log | show
echo "== clustered + reorder"; $B --clustered --reorder cluster 2>>gpurun_out/c_err.log | tee -a gpurun_out/c_variants.jsonl | show
M=l1tex__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum,sm__warps_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k regex:csr_spmm -s 8 -c 5 --csv --log-file gpurun_out/c_ncu_uniform.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-clustered --no-products --no-check > /dev/null 2>>gpurun_out/c_err.log
ncu --metrics $M --clock-control none -k regex:csr_spmm -s 8 -c 5 --csv --log-file gpurun_out/c_ncu_clustered_natural.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-clustered --no-products --no-check --clustered > /dev/null 2>>gpurun_out/c_err.log
ncu --metrics $M --clock-control none -k regex:csr_spmm -c 40 --csv --log-file gpurun_out/c_ncu_clustered_reorder.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-clustered --no-products --no-check --clustered --reorder cluster > /dev/null 2>>gpurun_out/c_err.log
tail -3 gpurun_out/c_err.log
