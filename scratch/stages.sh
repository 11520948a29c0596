#!/bin/bash
for cfg in cfg2 cfg5 cfg1; do for s in 2 3 4; do IMP_GPU_STAGES=$s python bench.py --config $cfg --steps 20 --e2e-steps 1 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$cfg stages=$s', round(d['value']), round(d['roofline']['frac'],4), round(d['ms_per_step'],4))"; done; done
