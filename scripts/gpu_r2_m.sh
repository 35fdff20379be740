#!/bin/bash
# Round 2, GPU job M: the lone-warp floor with kernel v14 (ncu --set full on the heaviest tile of pose 0 alone)
mkdir -p gpurun_out
for p in 0 1 3; do timeout 200 python scripts/lone_tile.py $p -1 8 2>/dev/null | tail -1 | tee gpurun_out/r2m_lone_v14_p$p.json; done
t=$(python -c "import json;print(json.load(open('gpurun_out/r2m_lone_v14_p0.json'))['tile'])")
timeout 400 ncu --set full --clock-control none --import-source on -k regex:render_fast --launch-skip 1 -c 1 -o gpurun_out/r2m_prof_lone_tile_v14 -f python scripts/lone_tile.py 0 $t 2 > gpurun_out/r2m_ncu_lone.log 2>&1; tail -2 gpurun_out/r2m_ncu_lone.log
