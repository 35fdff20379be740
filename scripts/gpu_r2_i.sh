#!/bin/bash
# Round 2, GPU job I: kernel v14 -- shard simulation, default bench, reference arm, launch list, ncu full capture, memcheck
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== shard simulation"
timeout 600 python scripts/shard_sim.py 0 1,2,4,8 2>gpurun_out/r2i_shard_sim.err | tee gpurun_out/r2i_shard_sim.jsonl; cp gpurun_out/shard_sim.json gpurun_out/r2i_shard_sim.json
echo "== bench default"
(time timeout 600 python bench.py) > gpurun_out/r2i_bench.json 2>gpurun_out/r2i_bench.err; tail -3 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],'e2e %.0f'%d['e2e']['value'],'frac %.3f'%d['roofline']['frac'],d['ms_per_frame_by_pose'])
e=d['extras']
for k in ('c3_2160p','c5_64_views','c4_dynamic_scene','moving_camera','tile_feedback_off','warm_l2'): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in e[k].items() if a not in ('note','rank0_stage_ms','frame_crc32_by_pose','path')})
PY
echo "== particles"
timeout 200 python scripts/particle_bench.py 1.0 1000000 0 2>/dev/null | tail -1 > gpurun_out/r2i_particles_fast.json; cut -c1-300 gpurun_out/r2i_particles_fast.json
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-c1 --no-extras > gpurun_out/r2i_launch_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/r2i_launches.csv | tail -12
echo "== ncu full"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_fast -c 4 -o gpurun_out/r2i_prof_v14 -f python scripts/profile_frame.py 1.0 4 > gpurun_out/r2i_ncu_full.log 2>&1; tail -2 gpurun_out/r2i_ncu_full.log
echo "== sanitizer (memcheck)"
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_group_gpu.py -m gpu -q -x -k "dynamic_tree_and_sparse or ragged or particle_step_equals_the_oracle or trace_lines or range_updates or growth or group_range_updates or gpu_tree_build" > gpurun_out/r2i_sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/r2i_sanitizer_memcheck.log
