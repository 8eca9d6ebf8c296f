#!/bin/bash
# evidence of the LAST build of the round: GPU test-suite, FLT32 CSR sweep under ncu (source of profiles/traffic.json),
# the default bench line + the reference arm on the same box, smoke
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gpu_tests.log 2>&1; tail -2 gpurun_out/r02c_gpu_tests.log
Q="--no-cpu --no-e2e --no-clustered --no-products --no-arxiv --no-check"
ncu --set full --clock-control none --import-source on -k regex:csr_spmm -s 5 -c 5 -o /tmp/ncu/r02c_csr -f python bench.py --steps 1 --warmup 1 $Q > /dev/null 2>>gpurun_out/r02c_err.log
python tools/ncu_summary.py /tmp/ncu/r02c_csr.ncu-rep > gpurun_out/r02c_csr_ncu_summary.json 2>>gpurun_out/r02c_err.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02c_bench_n1.json 2>> gpurun_out/r02c_err.log
python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/r02c_err.log | tail -1 > gpurun_out/r02c_bench_reference_arm.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>>gpurun_out/r02c_err.log
tail -3 gpurun_out/r02c_err.log
