#!/bin/bash
# Round 2, GPU job U (the last three GPU-minutes): v21 = the v19 experiment (-DQB_MASK_LATE) + the first ray's set-up by
# every lane (-DQB_FIRST_RAY_ALL_LANES), after which the SASS has no wait on the slot-record loads before the LOP3 that
# consumes them (scripts/sass_scoreboards.py).  A/B against v18 (the default build), lone tile, and -- only if v21 is
# ahead -- the GPU parity suite and a fuzz sweep on it.
mkdir -p gpurun_out
ab() { # lib tag n
  QB_CUC_LIB=$1 timeout 120 python bench.py --steps 24 --no-cpu --no-c1 --no-extras 2>gpurun_out/r2u_ab_$2_$3.err | tail -1 > gpurun_out/r2u_ab_$2_$3.json
  python - "$2" "$3" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/r2u_ab_%s_%s.json'%(sys.argv[1],sys.argv[2])))
    print(sys.argv[1],'ms/step %.4f'%d['ms_per_step'],{k:round(v,4) for k,v in d['ms_per_frame_by_pose'].items()},'crc',d['frame_crc32']['by_pose'],flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e,flush=True)
PY
}
lone() { QB_CUC_LIB=$PWD/ab/liboctree_cuc_$1.so timeout 60 python scripts/lone_tile.py $2 -1 8 2>/dev/null | tail -1 | tee gpurun_out/r2u_lone_$1_p$2.json; }
ab $PWD/ab/liboctree_cuc_v18.so v18 1
ab $PWD/ab/liboctree_cuc_v21.so v21 1
lone v21 0; lone v18 0
if python - <<'PY'
import json,sys
a=json.load(open('gpurun_out/r2u_ab_v18_1.json'))['ms_per_step']; b=json.load(open('gpurun_out/r2u_ab_v21_1.json'))['ms_per_step']
sys.exit(0 if b < a*0.998 else 1)
PY
then
  echo "== v21 ahead: pytest -m gpu on it"
  (time QB_CUC_LIB=$PWD/ab/liboctree_cuc_v21.so timeout 300 python -m pytest tests -m gpu -q -x) > gpurun_out/r2u_pytest_gpu_v21.log 2>&1; tail -4 gpurun_out/r2u_pytest_gpu_v21.log
  QB_CUC_LIB=$PWD/ab/liboctree_cuc_v21.so timeout 200 python scripts/parity_fuzz.py 100 40000 2>&1 | tail -1 | tee gpurun_out/r2u_parity_fuzz_v21.json
fi
ab $PWD/ab/liboctree_cuc_v18.so v18 2
ab $PWD/ab/liboctree_cuc_v21.so v21 2
lone v21 3; lone v18 3
