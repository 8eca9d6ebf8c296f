#!/bin/bash
# same-box A/B: a frozen baseline build (pygim_b200/libbackend_pim_base.so, e.g. built from HEAD) against the working
# tree's library - the pool's boxes differ by a few per cent, so variants are only compared inside one call
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_v3.py -m gpu -x -q -k "host or pipeline" 2>&1 | tail -2
i=0
for lib in libbackend_pim_base.so libbackend_pim.so libbackend_pim_base.so libbackend_pim.so; do
  i=$((i+1)); PYGIM_LIB_PATH=pygim_b200/$lib python bench.py --steps 10 --warmup 3 --no-cpu --no-clustered --no-products --no-arxiv > gpurun_out/v_$i.json 2>>gpurun_out/v_err.log
done
tail -3 gpurun_out/v_err.log
