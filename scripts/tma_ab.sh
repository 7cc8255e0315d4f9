#!/bin/bash
# A/B of the fused probe+aggregate kernel: register-staged (default) vs warp-specialised TMA ring (SQLRS_B200_TMA=1), Q3' SF10 and SF100.
# Per-launch device times from ncu's launch list (serialised, cold cache: compare the two variants, not absolutes) + wall time of the query.
set -e
for sf in 10 100; do
  for tma in 0 1; do
    if [ $tma = 1 ]; then export SQLRS_B200_TMA=1; else unset SQLRS_B200_TMA; fi
    echo "== SF$sf tma=$tma"
    python scripts/q3_time.py $sf 8 full 2>&1 | head -1
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"sq_joinagg" --csv --log-file gpurun_out/r02g_q3_sf${sf}_joinagg_tma${tma}.csv python scripts/q3_time.py $sf 3 full > /dev/null 2>&1
    python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02g_q3_sf${sf}_joinagg_tma${tma}.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); mi=h.index("Metric Name")
for r in rows[-4:]: print("   ", r[ki], r[mi], r[vi])
PY
  done
done
