#!/bin/bash
# Round 2, GPU job E: kernel v13 -- per-rank shares of an N-GPU split simulated on one GPU with the resident-CTA cap swept,
# full parity suite + fuzz, the bench line, ncu launch list and full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== shard simulation (occupancy cap sweep)"
timeout 600 python scripts/shard_sim.py 0,5,4,3,2 1,4,8 2>gpurun_out/r2e_shard_sim.err | tee gpurun_out/r2e_shard_sim.jsonl
echo "== pytest -m gpu"
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/r2e_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2e_pytest_gpu.log
echo "== parity fuzz"; timeout 300 python scripts/parity_fuzz.py 400 13000 > gpurun_out/r2e_fuzz.log 2>&1; tail -2 gpurun_out/r2e_fuzz.log
echo "== bench default"
(time timeout 600 python bench.py) > gpurun_out/r2e_bench.json 2>gpurun_out/r2e_bench.err; tail -c 300 gpurun_out/r2e_bench.json; tail -3 gpurun_out/r2e_bench.err
echo "== reference arm"
(time timeout 600 python bench.py --impl reference --steps 5 --warmup 1) > gpurun_out/r2e_bench_reference.json 2>gpurun_out/r2e_bench_reference.err; tail -c 600 gpurun_out/r2e_bench_reference.json
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-c1 --no-extras > gpurun_out/r2e_launch_bench.log 2>&1
echo "== ncu full"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_fast -c 4 -o gpurun_out/r2e_prof_v13 -f python scripts/profile_frame.py 1.0 4 > gpurun_out/r2e_ncu_full.log 2>&1; tail -2 gpurun_out/r2e_ncu_full.log
echo "== sanitizer (memcheck) on a small frame set"
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "dynamic_tree_and_sparse or ragged or particle_step_equals_the_oracle or trace_lines or range_updates" > gpurun_out/r2e_sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/r2e_sanitizer_memcheck.log
