#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --ignore=tests/test_gpu_fullsize.py > gpurun_out/i_tests.log 2>&1; echo "exit $?" >> gpurun_out/i_tests.log; tail -12 gpurun_out/i_tests.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/i_tests_full.log 2>&1; echo "exit $?" >> gpurun_out/i_tests_full.log; tail -4 gpurun_out/i_tests_full.log
python bench.py --steps 20 --warmup 3 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "bench exit $?"; tail -3 gpurun_out/i_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/i_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],2),'parity',d['parity_all_ranks'],d['parity_e2e'],'frac',round(d['roofline']['frac'],3),'cpu',d['cpu_baseline'] and round(d['cpu_baseline']['value']))
for p in d['per_hidden']: print('   H=%3d %.3f ms %.0f'%(p['hidden'],p['kernel_ms'],p['gflops']))
for name in ('clustered','products'):
    rec=d[name]
    if not rec: continue
    for k,v in rec.items():
        if isinstance(v,dict) and 'per_hidden' in v:
            print(name,k,round(v['value']),v.get('parity_all_ranks'),[round(p['kernel_ms'],3) for p in v['per_hidden']])
PY
for s in arxiv; do python bench.py --steps 50 --warmup 3 --no-cpu --no-e2e --no-clustered --no-products --shape $s 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('arxiv',[round(p['kernel_ms'],3) for p in d['per_hidden']])"; done
python examples/spmm_test.py --dataset Reddit --scale 0.2 --data_type INT32 --sp_format COO --hidden_size 64 --repeat 3 > gpurun_out/spmm_test_gpu_stdout.txt 2>gpurun_out/i_example.err; tail -4 gpurun_out/spmm_test_gpu_stdout.txt
