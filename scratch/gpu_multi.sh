#!/bin/bash
# multi-GPU round: product farm test on all GPUs, then the driver's torchrun bench line at N GPUs
N=${N:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_host_path.py -m gpu -q --tb=short -k "farm" 2>&1 | tail -5
SECONDS=0
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N wall: ${SECONDS}s rc=$?"; grep -E "Error|error|Traceback" gpurun_out/bench_n$N.err | head -5
python - <<PY
import json
try:
    l = json.loads([x for x in open("gpurun_out/bench_n$N.json").read().strip().splitlines() if x.startswith("{")][-1])
    print("HEAD n_gpus", l["n_gpus"], "ms/step %.4f value %.0f frac %.3f e2e %.0f (best %.0f worst %.0f)" % (l["ms_per_step"], l["value"], l["roofline"]["frac"], l["e2e"]["value"], l["e2e"]["best"], l["e2e"]["worst"]))
    print("h2d", l.get("h2d_ceiling"))
    for k, e in l.get("extra_configs", {}).items():
        print(k, "ms/step %.4f value %.0f frac(rank0) %.3f jobs/rank %d" % (e["ms_per_step"], e["value"], e["roofline"]["frac"], e["jobs_this_rank"]), "e2e", e.get("e2e", {}).get("value"))
        print("   size_aware", e.get("size_aware")); print("   farm", json.dumps(e.get("farm_api")))
except Exception as ex:
    print("FAILED", ex); print(open("gpurun_out/bench_n$N.err").read()[-3000:])
PY
