#!/bin/bash
# one GPU session: parity suite, then one device-timed bench line per config (no CPU leg)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short ${PYTEST_ARGS:--x} 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
for c in ${CONFIGS:-cfg2 cfg1 cfg1l cfg3 cfg3nn cfg4 cfg5}; do
  timeout 600 python bench.py --config $c --no-cpu --steps 20 --e2e-steps 3 > gpurun_out/b_$c.log 2> gpurun_out/b_$c.err
done
tail -3 gpurun_out/pytest_gpu.log
for c in ${CONFIGS:-cfg2 cfg1 cfg1l cfg3 cfg3nn cfg4 cfg5}; do python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/b_$c.log").read().strip().splitlines()[-1])
    print("$c", "ms/step %.4f"%l["ms_per_step"], "frac %.3f"%l["roofline"]["frac"], "e2e %.0f"%l["e2e"]["value"], "launches/step", l["roofline"]["kernel_launches_per_step"], "lat", l.get("single_request_latency_ms",{}).get("p50"))
except Exception as e:
    print("$c FAILED", e, open("gpurun_out/b_$c.err").read()[-600:])
PY
done
