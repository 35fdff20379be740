#!/bin/bash
# Round 2, GPU job R: the shipped kernel (v18) -- smoke, full GPU parity suite, fuzz sweep, the default bench line and
# the reference arm as the driver runs them, launch list, ncu --set full of the four poses, memcheck on a subset.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== pytest -m gpu"
(time timeout 900 python -m pytest tests -m gpu -q -x) > gpurun_out/r2r_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2r_pytest_gpu.log
echo "== parity fuzz"
timeout 600 python scripts/parity_fuzz.py 300 20000 2>&1 | tail -2 | tee gpurun_out/r2r_parity_fuzz.json
echo "== reference arm"
(time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2r_bench_reference.json 2>gpurun_out/r2r_bench_reference.err; cut -c1-400 gpurun_out/r2r_bench_reference.json
echo "== bench (driver's line)"
(time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2r_bench_driver_line.json 2>gpurun_out/r2r_bench_driver_line.err; tail -2 gpurun_out/r2r_bench_driver_line.err
echo "== bench default"
(time timeout 600 python bench.py) > gpurun_out/r2r_bench.json 2>gpurun_out/r2r_bench.err; tail -3 gpurun_out/r2r_bench.err
for f in gpurun_out/r2r_bench_driver_line.json gpurun_out/r2r_bench.json; do
python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],'e2e %.0f'%d['e2e']['value'],'frac %.3f'%d['roofline']['frac'],d['ms_per_frame_by_pose'], 'launches', d['gpu_launches'], 'cpu', d.get('cpu_baseline',{}).get('value'))
e=d.get('extras',{})
for k in ('c3_2160p','c5_64_views','c4_dynamic_scene','moving_camera','tile_feedback_off','warm_l2','c1_640x360'):
    if k in e: print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in e[k].items() if a not in ('note','rank0_stage_ms','frame_crc32_by_pose','path')})
PY
done
echo "== 4K"
timeout 300 python bench.py --res 4k --no-cpu --no-c1 --no-extras 2>/dev/null | tail -1 > gpurun_out/r2r_bench_4k.json; python -c "
import json; d=json.load(open('gpurun_out/r2r_bench_4k.json')); print('4k ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],'e2e %.0f'%d['e2e']['value'],'frac %.3f'%d['roofline']['frac'])"
echo "== lone tile"
for p in 0 1 3; do timeout 200 python scripts/lone_tile.py $p -1 8 2>/dev/null | tail -1 | tee gpurun_out/r2r_lone_v18_p$p.json; done
echo "== shard simulation"
timeout 300 python scripts/shard_sim.py 0 1,4,8 2>gpurun_out/r2r_shard_sim.err | tail -4 | cut -c1-400; cp gpurun_out/shard_sim.json gpurun_out/r2r_shard_sim.json 2>/dev/null
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-c1 --no-extras > gpurun_out/r2r_launch_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/r2r_launches.csv | tail -12
echo "== ncu full"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_fast -c 4 -o gpurun_out/r2r_prof_v18 -f python scripts/profile_frame.py 1.0 4 > gpurun_out/r2r_ncu_full.log 2>&1; tail -2 gpurun_out/r2r_ncu_full.log
echo "== sanitizer (memcheck)"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_group_gpu.py -m gpu -q -x -k "dynamic_tree_and_sparse or ragged or particle_step_equals_the_oracle or trace_lines or range_updates or growth or group_range_updates" > gpurun_out/r2r_sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/r2r_sanitizer_memcheck.log
