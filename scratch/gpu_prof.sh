#!/bin/bash
# ncu --set full of one kernel launch of each listed config (after the plain run exited 0); the report is condensed ON THE BOX
# (summary, per-opcode mix, gzipped source page with CUDA-C lines) and deleted: gpurun_out/ only carries 64 MiB back.
# usage: CONFIGS="cfg3 cfg4" TAG=r02a bash scratch/gpu_prof.sh
mkdir -p gpurun_out
TAG=${TAG:-r02}
for c in ${CONFIGS:-cfg3 cfg3nn cfg4}; do
  python bench.py --config $c --scale ${SCALE:-4} --steps 3 --e2e-steps 0 --no-cpu --extras none > gpurun_out/prof_$c.plain.log 2>&1 || { echo "$c plain run failed"; tail -5 gpurun_out/prof_$c.plain.log; continue; }
  ncu --set full --clock-control none --import-source on -k regex:'imp_(strip|blur_tile|cubic_tile|cubic_run|pass|gather_tile)_kernel' -s 2 -c 1 -f -o /tmp/${TAG}_$c \
      python bench.py --config $c --scale ${SCALE:-4} --steps 3 --e2e-steps 0 --no-cpu --extras none > gpurun_out/prof_$c.ncu.log 2>&1
  R=/tmp/${TAG}_$c.ncu-rep
  [ -f $R ] || { echo "$c: no report"; tail -5 gpurun_out/prof_$c.ncu.log; continue; }
  python profiles/ncu_summary.py $R > gpurun_out/${TAG}_$c.summary.txt
  ncu -i $R --page source --csv > /tmp/${TAG}_$c.src.csv 2>/dev/null
  python profiles/ncu_ops.py /tmp/${TAG}_$c.src.csv > gpurun_out/${TAG}_$c.opmix.txt
  ncu -i $R --page source --csv --print-source cuda,sass > /tmp/${TAG}_$c.srcc.csv 2>/dev/null || cp /tmp/${TAG}_$c.src.csv /tmp/${TAG}_$c.srcc.csv
  gzip -c /tmp/${TAG}_$c.srcc.csv > gpurun_out/${TAG}_$c.source.csv.gz
  ncu -i $R --page details --csv 2>/dev/null | gzip -c > gpurun_out/${TAG}_$c.details.csv.gz
  ls -la gpurun_out/${TAG}_$c.* | awk '{print $5, $9}'
  rm -f $R /tmp/${TAG}_$c.src*.csv
done
