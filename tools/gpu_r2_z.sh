#!/bin/bash
mkdir -p gpurun_out
for lib in libbackend_pim.so libbackend_pim_tdeep.so libbackend_pim.so libbackend_pim_tdeep.so; do
  echo "== $lib"; PYGIM_LIB_PATH=pygim_b200/$lib python tools/host_overhead_probe.py 2>&1 | grep "arxiv" | awk '{print "   ", $1, $2, $3, $4, $11, $12, $13, $14, $15}'
done
