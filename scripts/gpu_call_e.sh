#!/bin/bash
set -x
R=${1:-r01j}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "q3 or plan or smoke" 2>&1 | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for v in tma notma; do
  if [ $v = notma ]; then export SQLRS_B200_NO_TMA=1; fi
  timeout 200 python bench.py --query q3 --steps 20 --warmup 3 --cpu-rows 0 > gpurun_out/${R}_bench_q3_sf10_$v.json 2>> gpurun_out/${R}.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/${R}_bench_q3_sf10_$v.json | head -2
  timeout 200 python bench.py --query q3 --q3-sf 100 --steps 5 --warmup 3 --cpu-rows 0 > gpurun_out/${R}_bench_q3_sf100_$v.json 2>> gpurun_out/${R}.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/${R}_bench_q3_sf100_$v.json | head -2
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"sq_join" --log-file gpurun_out/${R}_q3_sf10_joinkernels_$v.csv python scripts/q3_time.py 10 2 > /dev/null 2>&1
  grep -E "sq_join" gpurun_out/${R}_q3_sf10_joinkernels_$v.csv | awk -F'","' '{print $5, $NF}' | tail -4
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"sq_join" --log-file gpurun_out/${R}_q3_sf100_joinkernels_$v.csv python scripts/q3_time.py 100 2 > /dev/null 2>&1
  grep -E "sq_join" gpurun_out/${R}_q3_sf100_joinkernels_$v.csv | awk -F'","' '{print $5, $NF}' | tail -4
done
tail -3 gpurun_out/${R}.err
