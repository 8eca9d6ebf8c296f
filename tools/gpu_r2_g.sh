#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:csr_spmm -s 1 -c 1 -o gpurun_out/g_arxiv_h32 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-clustered --no-products --no-check --shape arxiv --hidden 16 32 > /dev/null 2>>gpurun_out/g_err.log
ncu --set full --clock-control none --import-source on -k regex:csr_spmm -s 1 -c 1 -o gpurun_out/g_reddit_h32 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-clustered --no-products --no-check --hidden 16 32 > /dev/null 2>>gpurun_out/g_err.log
ncu --set full --clock-control none --import-source on -k regex:csr_spmm -s 5 -c 1 -o gpurun_out/g_clustered_h32 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-clustered --no-products --no-check --hidden 16 32 --clustered --reorder cluster > /dev/null 2>>gpurun_out/g_err.log
ncu --set full --clock-control none --import-source on -k regex:csr_hc -s 1 -c 1 -o gpurun_out/g_tiles_h32 -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-clustered --no-products --no-check --hidden 16 32 --clustered --reorder tiles --tile-super-nnz 131072 --hot-k 1024 > /dev/null 2>>gpurun_out/g_err.log
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/g_err.log
