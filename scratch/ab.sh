#!/bin/bash
# A/B inside one box: base .so vs current build, several workloads; prints value / frac / ms
for cfg in "$@"; do
  for lib in scratch/libimp_gpu_base.so ""; do
    IMP_GPU_LIB=$lib python bench.py --config $cfg --steps 20 --e2e-steps 1 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$cfg', '${lib:-new}', d['value'], d['roofline']['frac'], d['ms_per_step'])"
  done
done
