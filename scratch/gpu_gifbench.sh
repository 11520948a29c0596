#!/bin/bash
# cfg2 headline kept short; the point is extra_configs.cfg3/cfg3nn e2e vs e2e_gif_pages
python bench.py --steps 5 --e2e-steps 6 --no-cpu --extras cfg3,cfg3nn > gpurun_out/gifbench.json 2> gpurun_out/gifbench.err; echo rc=$?
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/gifbench.json') if l.startswith('{')][0])
for k,v in d['extra_configs'].items():
    g=v['e2e_gif_pages']
    print(k, 'e2e', round(v['e2e']['value']), 'h2d', v['e2e']['h2d_bytes_per_step'], '| gif pages', round(g['value']), 'best', round(g['best']), 'h2d', g['h2d_bytes_per_step'], 'launches', g['kernel_launches_per_step'])
    print('   thumbs', json.dumps(g['thumbnails']))
PY
tail -3 gpurun_out/gifbench.err
