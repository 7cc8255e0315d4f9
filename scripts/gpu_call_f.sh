#!/bin/bash
set -x
R=${1:-r01k}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --query q3 --steps 20 --warmup 3 --cpu-rows 0 > gpurun_out/${R}_bench_q3_sf10.json 2>> gpurun_out/${R}.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/${R}_bench_q3_sf10.json | head -2
timeout 200 python bench.py --query q3 --q3-sf 100 --steps 5 --warmup 3 --cpu-rows 0 > gpurun_out/${R}_bench_q3_sf100.json 2>> gpurun_out/${R}.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/${R}_bench_q3_sf100.json | head -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_q3_sf10_launches.csv python scripts/q3_time.py 10 2 > /dev/null 2>&1
tail -3 gpurun_out/${R}.err
