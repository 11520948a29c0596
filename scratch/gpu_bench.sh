#!/bin/bash
# the driver's own bench invocation (all extras, CPU baselines), timed
mkdir -p gpurun_out
SECONDS=0; python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench wall: ${SECONDS}s"; grep -E "Error|error|Traceback" gpurun_out/bench_default.err | head -5
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
    print("HEAD ms/step %.4f value %.0f frac %.3f e2e %.0f (best %.0f worst %.0f) launches/e2e-step %s" % (l["ms_per_step"], l["value"], l["roofline"]["frac"], l["e2e"]["value"], l["e2e"]["best"], l["e2e"]["worst"], l["e2e"]["kernel_launches_per_step"]))
    print("cpu", l.get("cpu_baseline"))
    print("h2d", l.get("h2d_ceiling"))
    print("lat", json.dumps(l.get("request_latency_ms"), indent=1))
    print("cache", l.get("plan_cache"))
    for k, e in l.get("extra_configs", {}).items():
        print(k, "ms/step %.4f value %.0f frac %.3f e2e %.0f launches/e2e-step %s" % (e["ms_per_step"], e["value"], e["roofline"]["frac"], e["e2e"]["value"], e["e2e"]["kernel_launches_per_step"]), "cpu", (e.get("cpu_baseline") or {}).get("value"), ((e.get("cpu_baseline") or {}).get("single_worker") or {}).get("value"))
except Exception as ex:
    print("FAILED", ex); print(open("gpurun_out/bench_default.err").read()[-3000:])
PY
