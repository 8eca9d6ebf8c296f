#!/bin/bash
for o in "seg_len=2048" "seg_len=4096" "seg_len=100000"; do
  echo "== opts: $o"; PROBE_OPTS="$o" python tools/host_overhead_probe.py 2>&1 | grep "reddit" | awk '{print "   ", $1, $2, $3, $4, $9, $10}'
done
show() { python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    print('   value %.0f GFLOP/s  parity %s' % (d['value'], d.get('parity_all_ranks')), [round(p['kernel_ms'],3) for p in d['per_hidden']])
"; }
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-clustered --no-products"
for sl in 1024 4096 8192; do echo "== reddit N=1 seg_len=$sl"; $B --opt seg_len=$sl 2>/dev/null | show; done
