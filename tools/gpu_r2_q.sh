#!/bin/bash
# same-box A/B: the tree at the start of this session (_old/) against the working tree
mkdir -p gpurun_out
for t in _old . _old .; do
  echo "== tree $t"; (cd $t && python tools/host_overhead_probe.py 2>&1 | grep "us" | awk '{print "   ", $1, $2, $3, $11, $12, $13, $14, $15}')
done
show() { python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print('   value %.0f GFLOP/s  parity %s' % (d['value'], d.get('parity_all_ranks')), [round(p['kernel_ms'],4) for p in d['per_hidden']])
"; }
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-clustered --no-products"
for t in _old . _old .; do echo "== reddit full, tree $t"; (cd $t && $B 2>/dev/null | show); done
