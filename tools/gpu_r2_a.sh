#!/bin/bash
# round-2 GPU call A: parity suite, sanitizer on smoke, probes, first bench line
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
./tools/l1_gather_probe > gpurun_out/a_l1_probe.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -x --ignore=tests/test_gpu_fullsize.py > gpurun_out/a_tests.log 2>&1
echo "exit $?" >> gpurun_out/a_tests.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --maxfail=25 > gpurun_out/a_tests_full.log 2>&1
echo "exit $?" >> gpurun_out/a_tests_full.log
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_sanitizer.log 2>&1
echo "exit $?" >> gpurun_out/a_sanitizer.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "exit $?" >> gpurun_out/a_bench.err
tail -5 gpurun_out/a_tests.log; tail -5 gpurun_out/a_tests_full.log; tail -3 gpurun_out/a_sanitizer.log; tail -3 gpurun_out/a_bench.err; head -c 600 gpurun_out/a_bench.json
