#!/bin/bash
# pageable host frames: batches and single requests, streaming copies on (default) / off
for nt in ${NTS:-1 0}; do echo "IMP_GPU_STAGE_NT=$nt (threads ${IMP_GPU_STAGE_THREADS:-default})"; IMP_GPU_STAGE_NT=$nt python scratch/pageable_ab.py 2>&1 | grep pageable
IMP_GPU_STAGE_NT=$nt python bench.py --steps 3 --e2e-steps 4 --no-cpu --extras cfg3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['request_latency_ms']
for k,v in r.items(): print(k, {a:(round(b['p50'],3) if isinstance(b,dict) else b) for a,b in v.items() if a!='what'})
g=d['extra_configs']['cfg3']['e2e_gif_pages']; print('gif pages', round(g['value']), g['thumbnails'])"
done
