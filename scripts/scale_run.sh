#!/bin/bash
# scripts/scale_run.sh N [extra bench args]: one torchrun bench at N GPUs, JSON to gpurun_out/scale_N_<tag>.json
N=$1; shift; TAG=$1; shift
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --no-cpu "$@" > gpurun_out/scale_${N}_${TAG}.json 2> gpurun_out/scale_${N}_${TAG}.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $N --no-cpu "$@" > gpurun_out/scale_${N}_${TAG}.json 2> gpurun_out/scale_${N}_${TAG}.err
fi
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_${N}_${TAG}.json"))
    print("N=${N} ${TAG}", "ms/step %.4f"%d["ms_per_step"], "Mrays/s %.0f"%d["value"], "e2e %.0f"%d["e2e"]["value"], {k:round(v,3) for k,v in d["ms_per_frame_by_pose"].items()})
except Exception as e:
    print("N=${N} ${TAG} FAILED", e); print(open("gpurun_out/scale_${N}_${TAG}.err").read()[-1500:])
PY
