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
timeout 600 python -m pytest tests/test_gpu_v3.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_boundary.py -m gpu -q --timeout 120 > gpurun_out/l_tests.log 2>&1; tail -3 gpurun_out/l_tests.log
for o in "" "item_nnz=128" "short_rows=3"; do
  echo "== opts: $o"; PROBE_OPTS="$o" python tools/host_overhead_probe.py 2>&1 | grep "per-call" | awk '{print "   ", $1, $2, $3, $4, $9, $10}'
done
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-clustered --no-products"
echo "== reddit"; $B 2>>gpurun_out/l_err.log | show
echo "== products"; $B --shape products 2>>gpurun_out/l_err.log | show
echo "== clustered cluster"; $B --clustered --reorder cluster 2>>gpurun_out/l_err.log | show
echo "== int8 coo"; $B --dtype INT8 --format COO 2>>gpurun_out/l_err.log | show
