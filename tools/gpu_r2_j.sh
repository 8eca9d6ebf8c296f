#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 --timeout 120 --ignore=tests/test_gpu_fullsize.py > gpurun_out/j_tests.log 2>&1; echo "exit $?" >> gpurun_out/j_tests.log; tail -12 gpurun_out/j_tests.log
for s in arxiv; do python bench.py --steps 50 --warmup 3 --no-cpu --no-e2e --no-clustered --no-products --shape $s 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('arxiv',[round(p['kernel_ms'],3) for p in d['per_hidden']], d['parity_all_ranks'])"; done
for w in "--dtype INT8 --format COO" "--dtype INT32 --format COO" "--dtype FLT32 --format COO" "--dtype INT8 --format CSR"; do
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-clustered --no-products $w 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', round(d['value']), [round(p['kernel_ms'],3) for p in d['per_hidden']], d['parity_all_ranks'])"; done
python bench.py --workload inference --steps 10 --dtype INT32 --format COO 2>/dev/null | tail -1 > gpurun_out/j_infer_i32coo.json; python -c "
import json; d=json.load(open('gpurun_out/j_infer_i32coo.json')); print('infer i32 coo', d['per_model'], d.get('cpu_baseline'))"
