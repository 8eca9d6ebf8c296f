#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/o_tests.log 2>&1; tail -3 gpurun_out/o_tests.log
echo "== default"; python tools/host_overhead_probe.py 2>&1 | grep "us"
echo "== short_rows=3"; PROBE_OPTS="short_rows=3" python tools/host_overhead_probe.py 2>&1 | grep "arxiv"
echo "== seg_len=128"; PROBE_OPTS="seg_len=128" python tools/host_overhead_probe.py 2>&1 | grep "arxiv"
echo "== seg_len=512"; PROBE_OPTS="seg_len=512" python tools/host_overhead_probe.py 2>&1 | grep "arxiv"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"csr_" -c 40 --csv --log-file gpurun_out/o_arxiv_launches.csv python bench.py --shape arxiv --steps 2 --warmup 1 --no-cpu --no-e2e --no-clustered --no-products --no-check > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/o_arxiv_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-12:]: print(r[4][:70], r[7], r[8], r[-1])
PY
