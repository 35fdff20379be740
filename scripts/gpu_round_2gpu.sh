#!/bin/bash
# 2-GPU job: the multi-GPU parity test, the strong-scaling bench at N = 1, 2 and the 4K line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > gpurun_out/pytest_2gpu.log 2>&1; tail -3 gpurun_out/pytest_2gpu.log
bash scripts/scale_run.sh 1 v11
bash scripts/scale_run.sh 2 v11
timeout 200 python bench.py --res 4k --no-cpu --no-c1 > gpurun_out/bench_4k_v11.json 2>gpurun_out/bench_4k_v11.err; python -c "
import json; d=json.load(open('gpurun_out/bench_4k_v11.json')); print('4k', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['ms_per_frame_by_pose'])"
