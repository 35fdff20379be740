#!/bin/bash
# Round 2, GPU job T: kernel v19 (slot records carry the child's mask replicated into the four bytes of the word, the
# two trees' masks are merged by the instruction that applies them at the end of the expansion -- nothing at the end
# of a descent waits for the record loads) -- A/B against v18 and against the sinking alone, parity suite, fuzz, the
# driver's bench line, lone tile, shard simulation, launch list, ncu, memcheck.  Ordered by importance: the job may be cut.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
ab() { # lib tag
  QB_CUC_LIB=$1 timeout 200 python bench.py --steps 24 --no-cpu --no-c1 --no-extras 2>gpurun_out/r2t_ab_$2.err | tail -1 > gpurun_out/r2t_ab_$2.json
  python - "$2" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/r2t_ab_%s.json'%sys.argv[1]))
    print(sys.argv[1],'ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],{k:round(v,4) for k,v in d['ms_per_frame_by_pose'].items()},'crc',d['frame_crc32']['by_pose'],flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e,flush=True)
PY
}
echo "== A/B full frame"
for v in v18 v19 v19plain v18 v19 v19plain; do ab $PWD/ab/liboctree_cuc_$v.so $v; cp gpurun_out/r2t_ab_$v.json gpurun_out/r2t_ab_${v}_$((++i)).json; done
echo "== pytest -m gpu"
(time timeout 600 python -m pytest tests -m gpu -q -x) > gpurun_out/r2t_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2t_pytest_gpu.log
echo "== parity fuzz (default build)"
timeout 300 python scripts/parity_fuzz.py 200 30000 2>&1 | tail -2 | tee gpurun_out/r2t_parity_fuzz.json
echo "== bench (driver's line)"
(time timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2t_bench_driver_line.json 2>gpurun_out/r2t_bench_driver_line.err; tail -2 gpurun_out/r2t_bench_driver_line.err
python - gpurun_out/r2t_bench_driver_line.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],'e2e %.0f'%d['e2e']['value'],'frac %.3f'%d['roofline']['frac'],d['ms_per_frame_by_pose'], 'launches', d['gpu_launches'], 'cpu', d.get('cpu_baseline',{}).get('value'))
e=d.get('extras',{})
for k in ('c3_2160p','c5_64_views','c4_dynamic_scene','moving_camera','tile_feedback_off','warm_l2','c1_640x360'):
    if k in e: print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in e[k].items() if a not in ('note','rank0_stage_ms','frame_crc32_by_pose','path')})
PY
echo "== lone tile (v19, then v18)"
for p in 0 3; do timeout 100 python scripts/lone_tile.py $p -1 8 2>/dev/null | tail -1 | tee gpurun_out/r2t_lone_v19_p$p.json; done
for p in 0 3; do QB_CUC_LIB=$PWD/ab/liboctree_cuc_v18.so timeout 100 python scripts/lone_tile.py $p -1 8 2>/dev/null | tail -1 | tee gpurun_out/r2t_lone_v18_p$p.json; done
echo "== shard simulation"
timeout 200 python scripts/shard_sim.py 0 1,4,8 2>gpurun_out/r2t_shard_sim.err | tail -4 | cut -c1-400; cp gpurun_out/shard_sim.json gpurun_out/r2t_shard_sim.json 2>/dev/null
echo "== ncu full"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_fast -c 4 -o gpurun_out/r2t_prof_v19 -f python scripts/profile_frame.py 1.0 4 > gpurun_out/r2t_ncu_full.log 2>&1; tail -2 gpurun_out/r2t_ncu_full.log
echo "== ncu launch list"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-c1 --no-extras > gpurun_out/r2t_launch_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/r2t_launches.csv | tail -12
echo "== 4K"
timeout 200 python bench.py --res 4k --no-cpu --no-c1 --no-extras 2>/dev/null | tail -1 > gpurun_out/r2t_bench_4k.json; python -c "
import json; d=json.load(open('gpurun_out/r2t_bench_4k.json')); print('4k ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],'e2e %.0f'%d['e2e']['value'],'frac %.3f'%d['roofline']['frac'])"
echo "== smoke"; timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== sanitizer (memcheck)"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_group_gpu.py -m gpu -q -x -k "dynamic_tree_and_sparse or ragged or particle_step_equals_the_oracle or trace_lines or range_updates or growth or group_range_updates" > gpurun_out/r2t_sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/r2t_sanitizer_memcheck.log
