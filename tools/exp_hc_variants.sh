#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
show() { python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print('   value %.0f GFLOP/s  parity %s' % (d['value'], d.get('parity_all_ranks')), [round(p['kernel_ms'],4) for p in d['per_hidden']], (d.get('reorder_stats') or {}).get('hot_coverage'))
"; }
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-clustered --no-products --clustered"
echo "== clustered natural"; $B 2>>gpurun_out/r_err.log | show
for lib in libbackend_pim.so libbackend_pim_hc1024.so libbackend_pim_hc768.so; do
  echo "== tiles $lib"; PYGIM_LIB_PATH=pygim_b200/$lib $B --reorder tiles 2>>gpurun_out/r_err.log | show
done
echo "== tiles hot_k 1024 super 131072"; $B --reorder tiles --tile-super-nnz 131072 --hot-k 1024 2>>gpurun_out/r_err.log | show
echo "== tiles hc1024 hot_k 640 super 32768"; PYGIM_LIB_PATH=pygim_b200/libbackend_pim_hc1024.so $B --reorder tiles --tile-super-nnz 32768 --hot-k 640 2>>gpurun_out/r_err.log | show
Q="--no-cpu --no-e2e --no-clustered --no-products --no-check"
ncu --set full --clock-control none --import-source on -k regex:csr_spmm_kernel -s 5 -c 2 -o /tmp/ncu/r_arxiv -f python bench.py --shape arxiv --steps 1 --warmup 1 $Q > /dev/null 2>>gpurun_out/r_err.log
cp /tmp/ncu/r_arxiv.ncu-rep gpurun_out/
tail -3 gpurun_out/r_err.log
