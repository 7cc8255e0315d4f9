#!/bin/bash
# One gpurun call: GPU parity tests, the driver's bench lines (both arms), the secondary Q3 lines, and the ncu
# evidence (launch lists + --set full of the dominant kernels).  Outputs land in gpurun_out/.
set -x
R=${1:-r01}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${R}_pytest_gpu.txt; cat gpurun_out/${R}_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${R}_bench_n1_sf100.json 2> gpurun_out/${R}_bench_n1.err; tail -2 gpurun_out/${R}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2>> gpurun_out/${R}_bench_n1.err
timeout 300 python bench.py --sf 10 --steps 30 --warmup 3 --cpu-rows 0 --e2e-steps 1 > gpurun_out/${R}_bench_n1_sf10.json 2>> gpurun_out/${R}_bench_n1.err
timeout 300 python bench.py --query q3 --steps 20 --warmup 3 > gpurun_out/${R}_bench_q3_sf10.json 2>> gpurun_out/${R}_bench_n1.err
timeout 300 python bench.py --query q3 --q3-sf 100 --steps 10 --warmup 3 --cpu-rows 0 > gpurun_out/${R}_bench_q3_sf100.json 2>> gpurun_out/${R}_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_q1_sf100_launches.csv python bench.py --sf 100 --steps 2 --warmup 1 --e2e-steps 0 --cpu-rows 0 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sq_agg_small -s 1 -c 1 -f -o gpurun_out/${R}_q1_small python bench.py --sf 100 --steps 2 --warmup 1 --e2e-steps 0 --cpu-rows 0 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_q3_sf10_launches.csv python scripts/q3_time.py 10 2 full > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_q3_sf100_launches.csv python scripts/q3_time.py 100 2 full > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sq_joinagg|sq_joinprobe|k_join_insert|k_join_probe_emit" -s 4 -c 4 -f -o gpurun_out/${R}_q3_kernels python scripts/q3_time.py 10 3 > /dev/null 2>&1
SQLRS_B200_TRACE=1 timeout 200 python scripts/q3_time.py 10 3 full 2>&1 | grep -E "trace|SF" | tail -28 > gpurun_out/${R}_q3_sf10_trace.txt
SQLRS_B200_TRACE=1 timeout 200 python scripts/q3_time.py 100 3 full 2>&1 | grep -E "trace|SF" | tail -28 > gpurun_out/${R}_q3_sf100_trace.txt
ls -la gpurun_out | tail -20
cut -c1-400 gpurun_out/${R}_bench_n1_sf100.json; cut -c1-300 gpurun_out/${R}_bench_q3_sf10.json; cut -c1-300 gpurun_out/${R}_bench_q3_sf100.json; tail -14 gpurun_out/${R}_q3_sf100_trace.txt
