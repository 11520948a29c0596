#!/bin/bash
# everything the driver runs at round end, plus the launch list of the same bench command
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash scratch/gpu_bench.sh
if [ -n "$LAUNCHES" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --e2e-steps 2 --no-cpu --extras none > gpurun_out/ncu_launches.log 2>&1
  python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); gi = hdr.index("Grid Size")
from collections import Counter
c = Counter(); t = Counter()
for r in rows[1:]:
    k = r[ki].split("(")[0]; c[(k, r[gi])] += 1; t[(k, r[gi])] += float(r[vi].replace(",", ""))
for k, n in c.most_common(12): print(n, k, "avg us %.1f" % (t[k] / n / 1000 if t[k] / n > 5000 else t[k] / n))
PY
fi
