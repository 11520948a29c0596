#!/bin/bash
# A/B inside one box: alternative builds (scratch/libimp_gpu_*.so) vs the current build; usage: scratch/ab.sh cfg4 cfg3 ...
for cfg in "$@"; do
  for lib in $(ls scratch/libimp_gpu_*.so 2>/dev/null) ""; do
    IMP_GPU_LIB=$lib python bench.py --config $cfg --steps 20 --e2e-steps 0 --no-cpu --extras none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$cfg', '${lib:-current}', 'ms %.4f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'])"
  done
done
