#!/bin/bash
# Round 2, GPU job O: kernel v16 (a pixel's results in three extra rows of the shared stack instead of registers, the
# camera ray evaluated again behind the loop: 72 -> 58 registers, 8 or 9 CTAs per SM) -- parity suite with the
# default build (8 CTAs), A/B over the resident-CTA targets, the default bench line with the threaded upload staging
# and the same with QB_UPLOAD_THREADS=0.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== pytest -m gpu"
(time timeout 900 python -m pytest tests -m gpu -q -x) > gpurun_out/r2o_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2o_pytest_gpu.log
ab() { # lib tag
  QB_CUC_LIB=$1 timeout 300 python bench.py --steps 24 --no-cpu --no-c1 --no-extras 2>gpurun_out/r2o_ab_$2.err | tail -1 > gpurun_out/r2o_ab_$2.json
  python - "$2" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/r2o_ab_%s.json'%sys.argv[1]))
    print(sys.argv[1],'ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],{k:round(v,4) for k,v in d['ms_per_frame_by_pose'].items()},'crc',d['frame_crc32']['by_pose'],flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e,flush=True)
PY
}
echo "== A/B full frame"
for v in v15 v16a v16b v16e v16f v16c v16d v15 v16b v16e; do ab $PWD/ab/liboctree_cuc_$v.so $v; done
echo "== parity fuzz (default build)"
timeout 600 python scripts/parity_fuzz.py 150 8000 2>&1 | tail -2 | tee gpurun_out/r2o_parity_fuzz.json
echo "== bench default"
(time timeout 600 python bench.py) > gpurun_out/r2o_bench.json 2>gpurun_out/r2o_bench.err; tail -3 gpurun_out/r2o_bench.err
show() {
python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],'e2e %.0f'%d['e2e']['value'],'frac %.3f'%d['roofline']['frac'],d['ms_per_frame_by_pose'])
e=d['extras']
for k in ('c3_2160p','c5_64_views','c4_dynamic_scene','moving_camera','tile_feedback_off','warm_l2'): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in e[k].items() if a not in ('note','rank0_stage_ms','frame_crc32_by_pose','path')})
PY
}
show gpurun_out/r2o_bench.json
echo "== bench default, QB_UPLOAD_THREADS=0"
(time QB_UPLOAD_THREADS=0 timeout 600 python bench.py --no-cpu --no-c1) > gpurun_out/r2o_bench_nostager.json 2>gpurun_out/r2o_bench_nostager.err; tail -3 gpurun_out/r2o_bench_nostager.err
show gpurun_out/r2o_bench_nostager.json
echo "== lone tile"
for p in 0 3; do timeout 200 python scripts/lone_tile.py $p -1 8 2>/dev/null | tail -1 | tee gpurun_out/r2o_lone_v16_p$p.json; done
for v in v16a v16e; do QB_CUC_LIB=$PWD/ab/liboctree_cuc_$v.so timeout 200 python scripts/lone_tile.py 0 -1 8 2>/dev/null | tail -1 | tee gpurun_out/r2o_lone_${v}_p0.json; done
echo "== ncu full"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_fast -c 4 -o gpurun_out/r2o_prof_v16 -f python scripts/profile_frame.py 1.0 4 > gpurun_out/r2o_ncu_full.log 2>&1; tail -2 gpurun_out/r2o_ncu_full.log
