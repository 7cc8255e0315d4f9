#!/bin/bash
set -x
R=${1:-r01}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${R}_bench_n1_sf100.json 2> gpurun_out/${R}_bench_n1.err; tail -2 gpurun_out/${R}_bench_n1.err
timeout 300 python bench.py --sf 10 --steps 30 --warmup 3 --cpu-rows 0 --e2e-steps 1 > gpurun_out/${R}_bench_n1_sf10.json 2>> gpurun_out/${R}_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_q1_sf100_launches.csv python bench.py --sf 100 --steps 2 --warmup 1 --e2e-steps 0 --cpu-rows 0 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sq_agg_small -s 1 -c 1 -f -o gpurun_out/${R}_q1_small python bench.py --sf 100 --steps 2 --warmup 1 --e2e-steps 0 --cpu-rows 0 > /dev/null 2>&1
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${R}_pytest_gpu.txt; cat gpurun_out/${R}_pytest_gpu.txt
cut -c1-300 gpurun_out/${R}_bench_n1_sf100.json
