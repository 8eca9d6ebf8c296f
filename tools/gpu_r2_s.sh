#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_v3.py tests/test_gpu_boundary.py -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-clustered --no-products"
i=0
for o in "host_chunks=4" "host_chunks=6" "host_tile_bytes=512"; do
  i=$((i+1)); echo "== $i: $o"; $B --opt $o > gpurun_out/s_$i.json 2>>gpurun_out/s_err.log
done
tail -3 gpurun_out/s_err.log
