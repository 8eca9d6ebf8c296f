#!/bin/bash
mkdir -p gpurun_out
show() { python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print('   value %.0f GFLOP/s  parity %s' % (d['value'], d.get('parity_all_ranks')), [round(p['kernel_ms'],3) for p in d['per_hidden']])
"; }
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-clustered --no-products"
timeout 600 python -m pytest tests/test_gpu_v3.py tests/test_gpu_parity.py -m gpu -q --timeout 120 -x > gpurun_out/k_tests.log 2>&1; tail -3 gpurun_out/k_tests.log
echo "== reddit"; $B 2>>gpurun_out/k_err.log | show
echo "== reddit super_nnz=48000"; $B --opt super_nnz=48000 2>>gpurun_out/k_err.log | show
echo "== reddit super_nnz=200000"; $B --opt super_nnz=200000 2>>gpurun_out/k_err.log | show
echo "== arxiv"; $B --shape arxiv --steps 50 2>>gpurun_out/k_err.log | show
echo "== products"; $B --shape products 2>>gpurun_out/k_err.log | show
echo "== clustered cluster"; $B --clustered --reorder cluster 2>>gpurun_out/k_err.log | show
tail -3 gpurun_out/k_err.log
