#!/bin/bash
# One GPU-box job for a new build of the hot kernel: A/B against the previous build, the GPU parity suite, a short
# random parity sweep, occupancy variants, and the ncu evidence.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
ab() { # lib tag
  QB_CUC_LIB=$1 timeout 300 python bench.py --steps 16 --no-cpu --no-c1 2>gpurun_out/ab_$2.err | tail -1 > gpurun_out/ab_$2.json
  python - "$2" <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/ab_%s.json'%sys.argv[1]))
    print(sys.argv[1],'ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],{k:round(v,3) for k,v in d['ms_per_frame_by_pose'].items()},'frac %.3f'%d['roofline']['frac'],flush=True)
except Exception as e: print(sys.argv[1],'FAILED',e,flush=True)
PY
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
ab $PWD/ab/liboctree_cuc_v10.so v10
ab $PWD/qubatron_b200/liboctree_cuc.so v11
echo "== pytest -m gpu"; (time timeout 420 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
echo "== parity fuzz"; timeout 150 python scripts/parity_fuzz.py 150 5000 > gpurun_out/fuzz.log 2>&1; tail -3 gpurun_out/fuzz.log
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-c1 > gpurun_out/launch_bench.log 2>&1
echo "== ncu full"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_fast -c 4 -o gpurun_out/prof_v11 -f python scripts/profile_frame.py 1.0 4 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
echo "== bench default"; timeout 400 python bench.py > gpurun_out/bench_default.json 2>gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json
ls -la gpurun_out
