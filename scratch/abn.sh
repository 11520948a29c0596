#!/bin/bash
# usage: abn.sh "cfgA cfgB" lib1 lib2 ...   ("new" = the in-tree build)
for cfg in $1; do
  for lib in "${@:2}"; do
    l=$lib; [ "$lib" = new ] && l=""
    IMP_GPU_LIB=$l python bench.py --config $cfg --steps 20 --e2e-steps 1 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$cfg', '$lib', round(d['value']), round(d['roofline']['frac'],4), round(d['ms_per_step'],4))"
  done
done
