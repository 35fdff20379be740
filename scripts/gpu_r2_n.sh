#!/bin/bash
# Round 2, GPU job N: kernel v15 (shading behind the loop, compact base-cube entry, 3-D launch grid) -- smoke, the
# full GPU parity suite and a fuzz sweep with the default build, then the variant A/B (ab/*.so built by
# scripts/build_variants.sh: v14 = all three switches restored, and each change alone), the default bench line,
# the launch list and one ncu --set full capture of the four poses.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== pytest -m gpu"
(time timeout 900 python -m pytest tests -m gpu -q -x) > gpurun_out/r2n_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2n_pytest_gpu.log
ab() { # lib tag
  QB_CUC_LIB=$1 timeout 300 python bench.py --steps 24 --no-cpu --no-c1 --no-extras 2>gpurun_out/r2n_ab_$2.err | tail -1 > gpurun_out/r2n_ab_$2.json
  python - "$2" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/r2n_ab_%s.json'%sys.argv[1]))
    print(sys.argv[1],'ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],{k:round(v,4) for k,v in d['ms_per_frame_by_pose'].items()},'crc',d['frame_crc32']['by_pose'],flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e,flush=True)
PY
}
echo "== A/B full frame"
for v in v14 v15 shade_only entry_only grid_only v14 v15; do ab $PWD/ab/liboctree_cuc_$v.so $v; done
echo "== parity fuzz (default build)"
timeout 600 python scripts/parity_fuzz.py 150 7000 2>&1 | tail -2 | tee gpurun_out/r2n_parity_fuzz.json
echo "== bench default"
(time timeout 600 python bench.py) > gpurun_out/r2n_bench.json 2>gpurun_out/r2n_bench.err; tail -3 gpurun_out/r2n_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],'e2e %.0f'%d['e2e']['value'],'frac %.3f'%d['roofline']['frac'],d['ms_per_frame_by_pose'])
e=d['extras']
for k in ('c3_2160p','c5_64_views','c4_dynamic_scene','moving_camera','tile_feedback_off','warm_l2'): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in e[k].items() if a not in ('note','rank0_stage_ms','frame_crc32_by_pose','path')})
PY
echo "== lone tile"
for p in 0 3; do timeout 200 python scripts/lone_tile.py $p -1 8 2>/dev/null | tail -1 | tee gpurun_out/r2n_lone_v15_p$p.json; done
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-c1 --no-extras > gpurun_out/r2n_launch_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/r2n_launches.csv | tail -12
echo "== ncu full"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_fast -c 4 -o gpurun_out/r2n_prof_v15 -f python scripts/profile_frame.py 1.0 4 > gpurun_out/r2n_ncu_full.log 2>&1; tail -2 gpurun_out/r2n_ncu_full.log
ls -la gpurun_out | tail -12
