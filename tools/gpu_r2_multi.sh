#!/bin/bash
# N-GPU call: NCCL tests + bench at N ranks (N = number of visible GPUs)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/m_tests_$N.log 2>&1; tail -5 gpurun_out/m_tests_$N.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 1200 $T bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/m_bench_$N.json 2> gpurun_out/m_bench_$N.err; echo "exit $?"
tail -3 gpurun_out/m_bench_$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/m_bench_$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', round(d['value']), 'no_exchange', d['no_exchange'] and round(d['no_exchange']['value']), 'e2e', d['e2e'] and round(d['e2e']['value']), 'parity', d['parity_all_ranks'], d['parity_e2e'], 'selftest', d['selftest_multi'])
for p in d['per_hidden']: print('   H=%3d %.3f ms'%(p['hidden'],p['kernel_ms']))
print('   no-exchange per H', d['no_exchange'] and [round(x,3) for x in d['no_exchange']['per_hidden_ms']])
pr=d['products']
if pr:
    s=pr['sharded']; print('products', round(s['value']), 'noex', round(s['no_exchange']['value']), 'floor_ms', round(s['exchange_floor_ms'],3), 'ms', round(s['ms_per_step'],3), 'n1', pr.get('n1_same_box'), 'speedup', pr.get('speedup_with_exchange'), pr.get('speedup_without_exchange'), 'parity', s.get('parity_all_ranks'))
PY
timeout 600 $T bench.py --gpus $N --steps 20 --warmup 3 --sync barrier --no-products --no-selftest --no-e2e > gpurun_out/m_bench_${N}_barrier.json 2>> gpurun_out/m_bench_$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/m_bench_${N}_barrier.json').read().strip().splitlines()[-1])
print('barrier sync: value', round(d['value']), [round(p['kernel_ms'],3) for p in d['per_hidden']])
PY
