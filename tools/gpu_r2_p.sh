#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
Q="--no-cpu --no-e2e --no-clustered --no-products --no-check"
ncu --set full --clock-control none --import-source on -k regex:csr_spmm_kernel -s 5 -c 2 -o /tmp/ncu/p_arxiv -f python bench.py --shape arxiv --steps 1 --warmup 1 $Q > /dev/null 2>gpurun_out/p_err.log
cp /tmp/ncu/p_arxiv.ncu-rep gpurun_out/
ls -la gpurun_out/p_arxiv.ncu-rep; tail -3 gpurun_out/p_err.log
