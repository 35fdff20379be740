#!/bin/bash
# Round 2, GPU job F: validation of the current build (warp-per-tile rank kernel, page-locked replication log, error
# handler): parity suite, default bench, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== pytest -m gpu"
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2f_pytest_gpu.log
echo "== bench default"
(time timeout 600 python bench.py) > gpurun_out/r2f_bench.json 2>gpurun_out/r2f_bench.err; tail -3 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print('ms/step %.4f'%d['ms_per_step'],'Mrays/s %.0f'%d['value'],'e2e %.0f'%d['e2e']['value'],'kernel mean %.4f'%d['roofline']['kernel_ms_mean'],d['ms_per_frame_by_pose'])
PY
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-c1 --no-extras > gpurun_out/r2f_launch_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/r2f_launches.csv | tail -8
