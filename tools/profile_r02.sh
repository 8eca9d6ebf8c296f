#!/bin/bash
# Round-2 evidence for profiles/ (1 GPU): launch list of the default bench command, full captures of the dominant
# kernels (FLT32 CSR sweep; INT8 / INT32 COO; clustered graph natural / reordered / tiles; products), probes.
# Only the JSON summaries travel back (gpurun_out is capped at 64 MiB); the FLT32 CSR report is kept whole.
mkdir -p gpurun_out /tmp/ncu
Q="--no-cpu --no-e2e --no-clustered --no-products --no-check"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"csr_|coo_|all_ones|quant|wait_flags" -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-clustered --no-products > gpurun_out/r02_launches_bench.log 2>&1
cap() { # name, kernel regex, skip, bench args...
  name=$1; k=$2; s=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 5 -o /tmp/ncu/$name -f python bench.py --steps 1 --warmup 1 $Q "$@" > /dev/null 2>>gpurun_out/p_err.log
  python tools/ncu_summary.py /tmp/ncu/$name.ncu-rep > gpurun_out/${name}_ncu_summary.json 2>>gpurun_out/p_err.log
}
cap r02_csr csr_spmm 5
cp /tmp/ncu/r02_csr.ncu-rep gpurun_out/
cap r02_coo_i8 csr_spmm 5 --dtype INT8 --format COO
cap r02_coo_i32 csr_spmm 5 --dtype INT32 --format COO
cap r02_clustered_natural csr_spmm 5 --clustered
cap r02_clustered_cluster csr_spmm 9 --clustered --reorder cluster
cap r02_clustered_tiles csr_hc 5 --clustered --reorder tiles --tile-super-nnz 131072 --hot-k 1024
cap r02_products csr_spmm 5 --shape products
cap r02_arxiv csr_spmm 5 --shape arxiv
./tools/l1_gather_probe > gpurun_out/r02_l1_gather_probe.txt 2>&1
./tools/l2_gather_probe > gpurun_out/r02_l2_gather_probe.txt 2>&1
timeout 900 python tools/cpu_protocol.py > gpurun_out/cpu_protocol.json 2> gpurun_out/cpu_protocol.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2>> gpurun_out/p_err.log
python bench.py --workload inference --steps 10 --dtype INT32 --format COO 2>>gpurun_out/p_err.log | tail -1 > gpurun_out/r02_infer_i32coo.json
python bench.py --workload inference --steps 10 --dtype FLT32 --format CSR 2>>gpurun_out/p_err.log | tail -1 > gpurun_out/r02_infer_f32csr.json
ls -la gpurun_out/; tail -3 gpurun_out/p_err.log; tail -2 gpurun_out/cpu_protocol.err
