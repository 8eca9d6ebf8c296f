#!/bin/bash
# N-GPU validation of the final build: NCCL tests + the default bench at N ranks
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/w_tests_$N.log 2>&1; tail -3 gpurun_out/w_tests_$N.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 1200 $T bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02b_bench_n$N.json 2> gpurun_out/w_bench_$N.err; echo "exit $?"
grep -v -i "warn\|OMP_NUM\|\*\*\*\*" gpurun_out/w_bench_$N.err | tail -3
