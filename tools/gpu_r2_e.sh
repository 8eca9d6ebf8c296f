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
timeout 900 python -m pytest tests/test_gpu_v3.py -m gpu -q -x -k "hot_cold or scheduling or autotuned" > gpurun_out/e_tests.log 2>&1; tail -15 gpurun_out/e_tests.log
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-clustered --no-products"
echo "== reddit uniform (new default)"; $B 2>>gpurun_out/e_err.log | tee -a gpurun_out/e.jsonl | show
echo "== products (new default light)"; $B --shape products 2>>gpurun_out/e_err.log | tee -a gpurun_out/e.jsonl | show
echo "== arxiv"; $B --shape arxiv 2>>gpurun_out/e_err.log | tee -a gpurun_out/e.jsonl | show
echo "== clustered + cluster"; $B --clustered --reorder cluster 2>>gpurun_out/e_err.log | tee -a gpurun_out/e.jsonl | show
for k in 1280 1536 1024; do
  echo "== clustered + tiles hot_k=$k"; $B --clustered --reorder tiles --hot-k $k 2>>gpurun_out/e_err.log | tee -a gpurun_out/e.jsonl | show
done
echo "== clustered + tiles super 131072"; $B --clustered --reorder tiles --tile-super-nnz 131072 2>>gpurun_out/e_err.log | tee -a gpurun_out/e.jsonl | show
echo "== uniform + tiles (no structure: what does it cost?)"; $B --reorder tiles 2>>gpurun_out/e_err.log | tee -a gpurun_out/e.jsonl | show
M=l1tex__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
ncu --metrics $M --clock-control none -k regex:csr_hc -c 10 --csv --log-file gpurun_out/e_ncu_tiles.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-clustered --no-products --no-check --clustered --reorder tiles > /dev/null 2>>gpurun_out/e_err.log
tail -5 gpurun_out/e_err.log
