#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/u_tests.log 2>&1; tail -3 gpurun_out/u_tests.log
echo "== default"; python tools/host_overhead_probe.py 2>&1 | grep "us" | awk '{print "   ", $1, $2, $3, $4, $11, $12, $13, $14, $15}'
python bench.py --steps 10 --warmup 3 --no-cpu --no-clustered --no-products > gpurun_out/u_bench.json 2>>gpurun_out/u_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"csr_" -c 40 --csv --log-file gpurun_out/u_arxiv_launches.csv python bench.py --shape arxiv --steps 2 --warmup 1 --no-cpu --no-e2e --no-clustered --no-products --no-check > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/u_arxiv_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-8:]: print(r[4][:70], r[7], r[8], r[-1])
PY
tail -3 gpurun_out/u_err.log
