#!/bin/bash
# pageable host frames: batches and single requests
python scratch/pageable_ab.py 2>&1 | grep pageable
python bench.py --steps 3 --e2e-steps 4 --no-cpu --extras none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['request_latency_ms']
for k,v in r.items(): print(k, {a:(round(b['p50'],3) if isinstance(b,dict) else b) for a,b in v.items() if a!='what'})"
timeout 600 python -m pytest tests/test_gpu_host_path.py -m gpu -q -x 2>&1 | tail -2
