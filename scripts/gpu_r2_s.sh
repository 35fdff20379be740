#!/bin/bash
# Round 2, GPU job S (gpurun --gpus 2): the shipped kernel (v18) on two real devices -- the group / torchrun parity
# tests, the C-ABI group bench and bench.py under the driver's torchrun launch line at N = 2.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv,noheader
echo "== multi-GPU parity tests on $N devices"
(time timeout 600 python -m pytest tests/test_group_gpu.py tests/test_multigpu_gpu.py -m gpu -q -x) > gpurun_out/r2s_pytest_multi_$N.log 2>&1; tail -5 gpurun_out/r2s_pytest_multi_$N.log
echo "== C ABI group bench (one process drives the GPUs)"
timeout 300 python scripts/group_bench.py $N 24 2>gpurun_out/r2s_group_$N.err | tee gpurun_out/r2s_group_$N.jsonl | cut -c1-500
tail -2 gpurun_out/r2s_group_$N.err
echo "== bench.py, the driver's launch line at N=$N"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $N --steps 40 --warmup 4 > gpurun_out/r2s_scale_$N.json 2> gpurun_out/r2s_scale_$N.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/r2s_scale_%s.json"%n).read().strip().splitlines()[-1])
    e=d.get("extras",{})
    print("N=%s"%n, "ms/step %.4f"%d["ms_per_step"], "Mrays/s %.0f"%d["value"], "e2e %.0f"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"],
          {k:round(v,3) for k,v in d["ms_per_frame_by_pose"].items()}, "crc", d["frame_crc32"]["by_pose"], d["frame_crc32"]["equal_unsharded_render_on_rank0"], d["frame_crc32"]["equal_cached_n1_run"])
    for k in ("c3_2160p","c5_64_views","c4_dynamic_scene"):
        if k in e: print("   ",k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in e[k].items() if a not in ("note","frame_crc32_by_pose","rank0_stage_ms")})
except Exception as ex:
    print("N=%s FAILED"%n, ex); print(open("gpurun_out/r2s_scale_%s.err"%n).read()[-2500:])
PY
