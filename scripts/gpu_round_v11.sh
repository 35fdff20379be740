#!/bin/bash
# Record of the v11 variant A/B: build variants of the hot kernel in ab/*.so, built with
# -DQB_V11_SELP / _STORE / _DYNSKIP / _SNAP / -DQB_MINBLOCKS=7 switches (a selp, b dynskip, j selp+snap, k selp+store,
# h selp+store+snap, i = h + 7 CTAs, f all, g = f + 7 CTAs), + the parity suite on two of them.  The winners
# (selp, store, 7 CTAs) are the source now and the switches are gone; to A/B a new idea, build it into ab/ with
# `make OUT=... EXTRA=-D...` and list it here.
mkdir -p gpurun_out
ab() { # lib tag
  QB_CUC_LIB=$1 timeout 300 python bench.py --steps 24 --no-cpu --no-c1 2>gpurun_out/ab_$2.err | tail -1 > gpurun_out/ab_$2.json
  python - "$2" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/ab_%s.json'%sys.argv[1]))
    print(sys.argv[1],'ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],{k:round(v,3) for k,v in d['ms_per_frame_by_pose'].items()},'frac %.3f'%d['roofline']['frac'],flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e,flush=True)
PY
}
for v in v10 v11a v11b v11j v11k v11h v11i v11f v11g v10; do ab $PWD/ab/liboctree_cuc_$v.so $v; done
for v in v11h v11f; do
  echo "== pytest -m gpu ($v)"; QB_CUC_LIB=$PWD/ab/liboctree_cuc_$v.so timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$v.log 2>&1; tail -2 gpurun_out/pytest_gpu_$v.log
  echo "== parity fuzz ($v)"; QB_CUC_LIB=$PWD/ab/liboctree_cuc_$v.so timeout 150 python scripts/parity_fuzz.py 150 7000 > gpurun_out/fuzz_$v.log 2>&1; tail -1 gpurun_out/fuzz_$v.log
done
