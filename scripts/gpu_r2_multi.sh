#!/bin/bash
# Round 2, multi-GPU job (gpurun --gpus N): the group / torchrun parity tests on real devices, the C-ABI group bench,
# and the driver's own launch line for bench.py at 1, 2, 4, .. N ranks.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv,noheader
nvidia-smi topo -m 2>/dev/null | head -12
echo "== multi-GPU parity tests on $N devices"
(time timeout 900 python -m pytest tests/test_group_gpu.py tests/test_multigpu_gpu.py -m gpu -q -x) > gpurun_out/r2m_pytest_multi_$N.log 2>&1; tail -6 gpurun_out/r2m_pytest_multi_$N.log
echo "== C ABI group bench (one process drives the GPUs)"
timeout 900 python scripts/group_bench.py $N 24 2>gpurun_out/r2m_group_$N.err | tee gpurun_out/r2m_group_$N.jsonl | cut -c1-600
tail -3 gpurun_out/r2m_group_$N.err
echo "== bench.py, the driver's launch line"
n=1
while [ $n -le $N ]; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 40 --warmup 4 > gpurun_out/r2m_scale_$n.json 2> gpurun_out/r2m_scale_$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $n --steps 40 --warmup 4 > gpurun_out/r2m_scale_$n.json 2> gpurun_out/r2m_scale_$n.err
  fi
  python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/r2m_scale_%s.json"%n).read().strip().splitlines()[-1])
    e=d.get("extras",{})
    print("N=%s"%n, "ms/step %.4f"%d["ms_per_step"], "Mrays/s %.0f"%d["value"], "e2e %.0f"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"],
          {k:round(v,3) for k,v in d["ms_per_frame_by_pose"].items()}, "crc", d["frame_crc32"]["equal_unsharded_render_on_rank0"], d["frame_crc32"]["equal_cached_n1_run"])
    for k in ("warm_l2","tile_feedback_off","l2_persisting_window","c3_2160p","c5_64_views","c4_dynamic_scene"):
        if k in e: print("   ",k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in e[k].items() if a not in ("note","frame_crc32_by_pose")})
except Exception as ex:
    print("N=%s FAILED"%n, ex); print(open("gpurun_out/r2m_scale_%s.err"%n).read()[-2500:])
PY
  n=$((n*2))
done
echo "== reference arm under torchrun at N=$N (thread count check)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --impl reference --gpus $N --steps 3 --warmup 1 --no-c1 > gpurun_out/r2m_ref_$N.json 2> gpurun_out/r2m_ref_$N.err
tail -c 700 gpurun_out/r2m_ref_$N.json
