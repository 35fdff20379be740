#!/bin/bash
# A/B the bench workload over differently tuned builds: scripts/ab.sh lib1.so lib2.so ...
for lib in "$@"; do
  QB_CUC_LIB=$PWD/$lib python bench.py --steps 16 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('$lib', 'ms/step %.4f'%d['ms_per_step'], 'Mrays/s %.0f'%d['value'], 'by pose', {k:round(v,3) for k,v in d['ms_per_frame_by_pose'].items()}, 'frac %.3f'%d['roofline']['frac'])
"
done
