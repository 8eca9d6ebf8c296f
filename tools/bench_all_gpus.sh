#!/bin/bash
# bench at N = number of visible GPUs; prints a digest
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 1200 $T bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "exit $?"
grep -v -i "warn\|OMP_NUM\|\*\*\*\*" gpurun_out/r02_bench_n$N.err | tail -5
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', round(d['value']), 'no_exchange', d['no_exchange'] and round(d['no_exchange']['value']), 'e2e', d['e2e'] and round(d['e2e']['value']), 'parity', d['parity_all_ranks'], d['parity_e2e'], 'selftest', d['selftest_multi'] and d['selftest_multi']['ok'])
print('   per H', [round(p['kernel_ms'],3) for p in d['per_hidden']], ' no-exchange', d['no_exchange'] and [round(x,3) for x in d['no_exchange']['per_hidden_ms']])
pr=d['products']
if pr:
    s=pr['sharded']; print('products', round(s['value']), 'noex', round(s['no_exchange']['value']), 'floor_ms', round(s['exchange_floor_ms'],3), 'ms', round(s['ms_per_step'],3), [round(p['kernel_ms'],3) for p in s['per_hidden']], 'noex per H', [round(x,3) for x in s['no_exchange']['per_hidden_ms']], 'n1', pr.get('n1_same_box'), 'speedup', pr.get('speedup_with_exchange'), pr.get('speedup_without_exchange'), 'parity', s.get('parity_all_ranks'))
PY
