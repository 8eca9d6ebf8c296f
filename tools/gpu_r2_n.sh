#!/bin/bash
# round 2, session 2: GPU tests of the batch host pipeline + tiny-row launch, arxiv / products family timing, full bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/n_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/n_tests.log
show() { python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print('   value %.0f GFLOP/s  parity %s' % (d['value'], d.get('parity_all_ranks')), [round(p['kernel_ms'],4) for p in d['per_hidden']], 'e2e', d.get('e2e') and (round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3)))
"; }
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-clustered --no-products"
for o in "short_rows=4" "short_rows=3" "short_rows=2"; do echo "== arxiv $o"; $B --shape arxiv --opt $o 2>>gpurun_out/n_err.log | show; done
for o in "short_rows=3" "short_rows=4"; do echo "== products $o"; $B --shape products --opt $o 2>>gpurun_out/n_err.log | show; done
echo "== full bench"
python bench.py --steps 20 --warmup 3 > gpurun_out/n_bench.json 2>>gpurun_out/n_err.log; cat gpurun_out/n_bench.json | show
tail -5 gpurun_out/n_err.log
