#!/bin/bash
# same-box A/B: a frozen baseline build (pygim_b200/libbackend_pim_base.so, e.g. built from HEAD) against the working
# tree's library - the pool's boxes differ by a few per cent, so variants are only compared inside one call
mkdir -p gpurun_out
for lib in libbackend_pim_base.so libbackend_pim.so libbackend_pim_base.so libbackend_pim.so; do
  echo "== $lib"; PYGIM_LIB_PATH=pygim_b200/$lib python tools/host_overhead_probe.py 2>&1 | grep "us" | awk '{print "   ", $1, $2, $3, $4, $11, $12, $13, $14, $15}'
done
i=0
for lib in libbackend_pim_base.so libbackend_pim.so libbackend_pim_base.so libbackend_pim.so; do
  i=$((i+1)); PYGIM_LIB_PATH=pygim_b200/$lib python bench.py --steps 20 --warmup 3 --no-cpu --no-clustered --no-products --no-arxiv > gpurun_out/v_$i.json 2>>gpurun_out/v_err.log
done
tail -3 gpurun_out/v_err.log
