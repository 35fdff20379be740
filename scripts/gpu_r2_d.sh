#!/bin/bash
# Round 2, GPU job D: packed-fp32 (FADD2 / FMUL2) variant of the traversal arithmetic: A/B, parity suite and fuzz on the variant
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
ab() { # lib tag
  QB_CUC_LIB=$1 timeout 300 python bench.py --steps 24 --no-cpu --no-c1 --no-extras 2>gpurun_out/r2d_ab_$2.err | tail -1 > gpurun_out/r2d_ab_$2.json
  python - "$2" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/r2d_ab_%s.json'%sys.argv[1]))
    print(sys.argv[1],'ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],{k:round(v,3) for k,v in d['ms_per_frame_by_pose'].items()},flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e,flush=True)
PY
}
for v in "$@"; do ab $PWD/ab/liboctree_cuc_$v.so $v; done
ab $PWD/ab/liboctree_cuc_base.so base_again
last="${@: -1}"
echo "== pytest -m gpu on $last"
(time QB_CUC_LIB=$PWD/ab/liboctree_cuc_$last.so timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/r2d_pytest_$last.log 2>&1; tail -5 gpurun_out/r2d_pytest_$last.log
echo "== parity fuzz on $last"; QB_CUC_LIB=$PWD/ab/liboctree_cuc_$last.so timeout 300 python scripts/parity_fuzz.py 300 11000 > gpurun_out/r2d_fuzz_$last.log 2>&1; tail -2 gpurun_out/r2d_fuzz_$last.log
for p in 0 3; do QB_CUC_LIB=$PWD/ab/liboctree_cuc_$last.so timeout 200 python scripts/lone_tile.py $p -1 8 2>/dev/null | tail -1 | tee gpurun_out/r2d_lone_${last}_p$p.json; done
