#!/bin/bash
# Round 2, first GPU job: the GPU parity suite (with the new multi-GPU-behind-the-C-ABI tests, which run on one GPU
# as several shards) and the default bench line.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== new group tests first"
(time timeout 600 python -m pytest tests/test_group_gpu.py -x -q) > gpurun_out/r2a_pytest_group.log 2>&1; tail -15 gpurun_out/r2a_pytest_group.log
echo "== pytest -m gpu"
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r2a_pytest_gpu.log 2>&1; tail -8 gpurun_out/r2a_pytest_gpu.log
echo "== bench default"
(time timeout 600 python bench.py --steps 20 --warmup 5) > gpurun_out/r2a_bench.json 2>gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
