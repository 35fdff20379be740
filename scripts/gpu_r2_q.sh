#!/bin/bash
# Round 2, GPU job Q: kernel v18 (root mask in shared memory, tile divisions by multiplication, one range test for the
# three reciprocals of a ray, cube corner moved by fma with the octant bits from a shared table) -- parity suite, fuzz,
# A/B against v17 and against v18 with the corner selects, bench, ncu.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== pytest -m gpu"
(time timeout 900 python -m pytest tests -m gpu -q -x) > gpurun_out/r2q_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2q_pytest_gpu.log
ab() { # lib tag
  QB_CUC_LIB=$1 timeout 300 python bench.py --steps 24 --no-cpu --no-c1 --no-extras 2>gpurun_out/r2q_ab_$2.err | tail -1 > gpurun_out/r2q_ab_$2.json
  python - "$2" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/r2q_ab_%s.json'%sys.argv[1]))
    print(sys.argv[1],'ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],{k:round(v,4) for k,v in d['ms_per_frame_by_pose'].items()},'crc',d['frame_crc32']['by_pose'],flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e,flush=True)
PY
}
echo "== A/B full frame"
for v in v17 v18 v18sel v17 v18 v18sel; do ab $PWD/ab/liboctree_cuc_$v.so $v; done
echo "== parity fuzz (default build)"
timeout 600 python scripts/parity_fuzz.py 200 10000 2>&1 | tail -2 | tee gpurun_out/r2q_parity_fuzz.json
echo "== bench default"
(time timeout 600 python bench.py) > gpurun_out/r2q_bench.json 2>gpurun_out/r2q_bench.err; tail -3 gpurun_out/r2q_bench.err
show() {
python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],'e2e %.0f'%d['e2e']['value'],'frac %.3f'%d['roofline']['frac'],d['ms_per_frame_by_pose'])
e=d['extras']
for k in ('c3_2160p','c5_64_views','c4_dynamic_scene','moving_camera','tile_feedback_off','warm_l2'): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in e[k].items() if a not in ('note','rank0_stage_ms','frame_crc32_by_pose','path')})
PY
}
show gpurun_out/r2q_bench.json
for t in; do
echo "== dynamic scene, QB_UPLOAD_THREADS=$t"
QB_UPLOAD_THREADS=$t timeout 600 python bench.py --no-cpu --no-c1 --steps 8 > gpurun_out/r2q_bench_ut$t.json 2>gpurun_out/r2q_bench_ut$t.err; grep -i upload gpurun_out/r2q_bench_ut$t.err
python - $t <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2q_bench_ut%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
c=d['extras']['c4_dynamic_scene']; print({k:round(c[k],3) for k in ('range_upload_calls_ms','frame_ms','skin_and_build_ms','render_ms')})
PY
done
echo "== lone tile"
for p in 0 3; do timeout 200 python scripts/lone_tile.py $p -1 8 2>/dev/null | tail -1 | tee gpurun_out/r2q_lone_v18_p$p.json; done
echo "== particles"
timeout 200 python scripts/particle_bench.py 1.0 1000000 0 2>/dev/null | tail -1 > gpurun_out/r2q_particles.json; cut -c1-300 gpurun_out/r2q_particles.json
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-c1 --no-extras > gpurun_out/r2q_launch_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/r2q_launches.csv | tail -12
echo "== ncu full"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_fast -c 4 -o gpurun_out/r2q_prof_v18 -f python scripts/profile_frame.py 1.0 4 > gpurun_out/r2q_ncu_full.log 2>&1; tail -2 gpurun_out/r2q_ncu_full.log
