#!/bin/bash
set -x
R=${1:-r01f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${R}_pytest_gpu.txt; cat gpurun_out/${R}_pytest_gpu.txt
timeout 300 python bench.py --query q3 --steps 10 --warmup 3 --cpu-rows 0 > gpurun_out/${R}_bench_q3_sf10.json 2> gpurun_out/${R}_bench_q3.err; tail -3 gpurun_out/${R}_bench_q3.err; cut -c1-300 gpurun_out/${R}_bench_q3_sf10.json; grep -o '"full_query": {"ms_per_step": [0-9.]*' gpurun_out/${R}_bench_q3_sf10.json
SQLRS_B200_TRACE=1 timeout 200 python scripts/q3_time.py 10 3 full 2>&1 | grep -E "trace|SF" | tail -28 > gpurun_out/${R}_q3_full_trace.txt; cat gpurun_out/${R}_q3_full_trace.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_q3_sf10_launches.csv python scripts/q3_time.py 10 2 full > /dev/null 2>&1
timeout 300 python bench.py --query q3 --q3-sf 100 --steps 5 --warmup 3 --cpu-rows 0 > gpurun_out/${R}_bench_q3_sf100.json 2>> gpurun_out/${R}_bench_q3.err; cut -c1-300 gpurun_out/${R}_bench_q3_sf100.json; grep -o '"full_query": {"ms_per_step": [0-9.]*' gpurun_out/${R}_bench_q3_sf100.json
SQLRS_B200_TRACE=1 timeout 300 python scripts/q3_time.py 100 3 2>&1 | grep -E "trace|SF" | tail -24 > gpurun_out/${R}_q3_sf100_trace.txt; cat gpurun_out/${R}_q3_sf100_trace.txt
tail -5 gpurun_out/${R}_bench_q3.err
