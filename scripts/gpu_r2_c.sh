#!/bin/bash
# Round 2, GPU job C: the no-clamp kernel + fast single-ray traversal (particles, trace_lines): parity suite, fuzz,
# bench line, particle bench (fast vs generic), L2 window on the lone tile, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== pytest -m gpu"
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/r2c_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2c_pytest_gpu.log
echo "== parity fuzz"; timeout 240 python scripts/parity_fuzz.py 200 9000 > gpurun_out/r2c_fuzz.log 2>&1; tail -2 gpurun_out/r2c_fuzz.log
echo "== bench default"
(time timeout 600 python bench.py) > gpurun_out/r2c_bench.json 2>gpurun_out/r2c_bench.err; tail -c 400 gpurun_out/r2c_bench.json; tail -3 gpurun_out/r2c_bench.err
echo "== particles / trace_lines: fast vs generic traversal"
timeout 200 python scripts/particle_bench.py 1.0 1000000 0 2>/dev/null | tail -1 > gpurun_out/r2c_particles_fast.json; cut -c1-700 gpurun_out/r2c_particles_fast.json
timeout 200 python scripts/particle_bench.py 1.0 1000000 1 2>/dev/null | tail -1 > gpurun_out/r2c_particles_generic.json; cut -c1-700 gpurun_out/r2c_particles_generic.json
echo "== lone tile: L2 persisting window"
for p in 0 3; do
  timeout 200 python scripts/lone_tile.py $p -1 8 2>/dev/null | tail -1 | tee gpurun_out/r2c_lone_p$p.json
  t=$(python -c "import json;print(json.load(open('gpurun_out/r2c_lone_p$p.json'))['tile'])")
  QB_L2_WINDOW_MB=96 timeout 200 python scripts/lone_tile.py $p $t 8 2>gpurun_out/r2c_lone_win_p$p.err | tail -1 | tee gpurun_out/r2c_lone_win_p$p.json
  grep "L2 window" gpurun_out/r2c_lone_win_p$p.err | head -1
done
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-c1 --no-extras > gpurun_out/r2c_launch_bench.log 2>&1
echo "== ncu full"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_fast -c 4 -o gpurun_out/r2c_prof_v12 -f python scripts/profile_frame.py 1.0 4 > gpurun_out/r2c_ncu_full.log 2>&1; tail -2 gpurun_out/r2c_ncu_full.log
ls -la gpurun_out | tail -14
