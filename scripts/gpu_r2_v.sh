#!/bin/bash
# Round 2, GPU job V (the last GPU-minute and a half): the default build is now v21 -- the GPU parity suite on it, then
# a short fuzz sweep and the shard simulation, in that order (the job may be cut).
mkdir -p gpurun_out
(time timeout 80 python -m pytest tests -m gpu -q -x) > gpurun_out/r2v_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2v_pytest_gpu.log
timeout 40 python scripts/parity_fuzz.py 60 50000 2>&1 | tail -1 | tee gpurun_out/r2v_parity_fuzz.json
timeout 40 python scripts/shard_sim.py 0 1,8 2>/dev/null | tail -3 | cut -c1-300; cp gpurun_out/shard_sim.json gpurun_out/r2v_shard_sim.json 2>/dev/null
