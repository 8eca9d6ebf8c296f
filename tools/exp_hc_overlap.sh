#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_v3.py -m gpu -x -q -k "hot_cold or row_map" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "clustered" 2>&1 | tail -2
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-clustered --no-products --no-arxiv --clustered"
i=0
for o in "--reorder tiles" "--reorder tiles --opt hc_overlap=0" "--reorder tiles --tile-super-nnz 131072 --hot-k 1024" "--reorder tiles --tile-super-nnz 131072 --hot-k 1024 --opt hc_overlap=0" ""; do
  i=$((i+1)); echo "== $i: $o"; $B $o > gpurun_out/x_$i.json 2>>gpurun_out/x_err.log
done
tail -3 gpurun_out/x_err.log
