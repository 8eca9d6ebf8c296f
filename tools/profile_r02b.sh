#!/bin/bash
# Round-2 evidence of the FINAL build (1 GPU): full GPU test-suite, launch list of the default bench command, full
# captures of the FLT32 CSR sweep and of the arxiv-shape two-launch family, the default bench line, config-3/4 lines.
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gpu_tests.log 2>&1; tail -2 gpurun_out/r02b_gpu_tests.log
Q="--no-cpu --no-e2e --no-clustered --no-products --no-arxiv --no-check"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"csr_|coo_|all_ones|quant|wait_flags" -c 400 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-clustered --no-products --no-arxiv > gpurun_out/r02b_launches_bench.log 2>&1
cap() { # name, kernel regex, skip, count, bench args...
  name=$1; k=$2; s=$3; c=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -o /tmp/ncu/$name -f python bench.py --steps 1 --warmup 1 $Q "$@" > /dev/null 2>>gpurun_out/r02b_err.log
  python tools/ncu_summary.py /tmp/ncu/$name.ncu-rep > gpurun_out/${name}_ncu_summary.json 2>>gpurun_out/r02b_err.log
}
cap r02b_csr csr_spmm 5 5
cap r02b_arxiv "csr_" 10 10 --shape arxiv
python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench_n1.json 2>> gpurun_out/r02b_err.log
python bench.py --steps 10 --warmup 3 --dtype INT8 --format COO --no-cpu --no-clustered --no-products --no-arxiv 2>>gpurun_out/r02b_err.log | tail -1 > gpurun_out/r02b_bench_i8coo.json
python bench.py --steps 10 --warmup 3 --dtype INT32 --format COO --no-cpu --no-clustered --no-products --no-arxiv 2>>gpurun_out/r02b_err.log | tail -1 > gpurun_out/r02b_bench_i32coo.json
python bench.py --workload inference --steps 10 --dtype INT32 --format COO 2>>gpurun_out/r02b_err.log | tail -1 > gpurun_out/r02b_infer_i32coo.json
python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/r02b_err.log | tail -1 > gpurun_out/r02b_bench_reference_arm.json
ls -la gpurun_out/r02b_*; tail -3 gpurun_out/r02b_err.log
echo "== products family 4 vs default"
for o in "short_rows=3" "short_rows=4"; do python bench.py --shape products --steps 10 --warmup 3 --no-cpu --no-e2e --no-clustered --no-products --no-arxiv --opt $o 2>>gpurun_out/r02b_err.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$o', round(d['value']), [round(p['kernel_ms'],3) for p in d['per_hidden']], d['parity_all_ranks'])"; done
