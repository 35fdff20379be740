#!/bin/bash
# Round 2, GPU job B: the group fix + the new 4K / C5 parity tests, then the kernel variant A/B (ab/*.so, built by
# scripts/build_variants.sh): full-frame bench, lanes per instruction (ncu), and the lone-tile floor with its ncu capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== pytest -m gpu"
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/r2b_pytest_gpu.log 2>&1; tail -6 gpurun_out/r2b_pytest_gpu.log
ab() { # lib tag [env]
  env $3 QB_CUC_LIB=$1 timeout 300 python bench.py --steps 24 --no-cpu --no-c1 --no-extras 2>gpurun_out/r2b_ab_$2.err | tail -1 > gpurun_out/r2b_ab_$2.json
  python - "$2" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/r2b_ab_%s.json'%sys.argv[1]))
    print(sys.argv[1],'ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],{k:round(v,3) for k,v in d['ms_per_frame_by_pose'].items()},flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e,flush=True)
PY
}
echo "== A/B full frame"
for v in base noclamp prefetch ballot4 ballot8 ballot16 base; do ab $PWD/ab/liboctree_cuc_$v.so $v QB_X=0; done
ab $PWD/ab/liboctree_cuc_prefetch.so prefetch_on QB_PREFETCH=1
ab $PWD/ab/liboctree_cuc_noclamp_pf.so noclamp_pf_on QB_PREFETCH=1
echo "== lone tile (pose 0 and 3)"
for p in 0 3; do
  QB_CUC_LIB=$PWD/ab/liboctree_cuc_base.so timeout 200 python scripts/lone_tile.py $p -1 8 2>/dev/null | tail -1 | tee gpurun_out/r2b_lone_base_p$p.json
  t=$(python -c "import json;print(json.load(open('gpurun_out/r2b_lone_base_p$p.json'))['tile'])")
  QB_PREFETCH=1 QB_CUC_LIB=$PWD/ab/liboctree_cuc_prefetch.so timeout 200 python scripts/lone_tile.py $p $t 8 2>/dev/null | tail -1 | tee gpurun_out/r2b_lone_prefetch_p$p.json
  QB_CUC_LIB=$PWD/ab/liboctree_cuc_noclamp.so timeout 200 python scripts/lone_tile.py $p $t 8 2>/dev/null | tail -1 | tee gpurun_out/r2b_lone_noclamp_p$p.json
done
echo "== ncu: lanes per instruction, base vs ballot"
for v in base ballot4 ballot8 ballot16; do
  QB_CUC_LIB=$PWD/ab/liboctree_cuc_$v.so timeout 300 ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:render_fast -c 4 --csv --log-file gpurun_out/r2b_lanes_$v.csv python scripts/profile_frame.py 1.0 4 > /dev/null 2>&1
  python - $v <<'PY'
import csv,sys
rows=[r for r in csv.reader(open('gpurun_out/r2b_lanes_%s.csv'%sys.argv[1])) if len(r)>10]
h=rows[0]; out={}
for r in rows[1:]:
    d=dict(zip(h,r)); out.setdefault(d['ID'],{})[d['Metric Name']]=d['Metric Value']
print(sys.argv[1], out)
PY
done
echo "== ncu full: the heaviest tile of pose 0 alone (lone-warp floor)"
t=$(python -c "import json;print(json.load(open('gpurun_out/r2b_lone_base_p0.json'))['tile'])")
QB_CUC_LIB=$PWD/ab/liboctree_cuc_base.so timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_fast --launch-skip 1 -c 1 -o gpurun_out/r2b_prof_lone_tile -f python scripts/lone_tile.py 0 $t 2 > gpurun_out/r2b_ncu_lone.log 2>&1; tail -2 gpurun_out/r2b_ncu_lone.log
ls -la gpurun_out | tail -30
