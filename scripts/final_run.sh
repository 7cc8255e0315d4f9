#!/bin/bash
# One gpurun call: GPU parity tests, the driver's bench lines (both arms), the secondary Q3 line, and the ncu
# evidence (launch list + --set full of the dominant kernels).  Outputs land in gpurun_out/.
set -x
R=${1:-r01}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${R}_pytest_gpu.txt; cat gpurun_out/${R}_pytest_gpu.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/${R}_bench_n1_sf100.json 2> gpurun_out/${R}_bench_n1.err; tail -2 gpurun_out/${R}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2>> gpurun_out/${R}_bench_n1.err
python bench.py --sf 10 --steps 30 --warmup 3 --cpu-rows 0 --e2e-steps 1 > gpurun_out/${R}_bench_n1_sf10.json 2>> gpurun_out/${R}_bench_n1.err
python bench.py --query q3 --steps 10 --warmup 3 > gpurun_out/${R}_bench_q3_sf10.json 2>> gpurun_out/${R}_bench_n1.err
python scripts/q1_variants.py 10 > gpurun_out/${R}_q1_sf10_variants.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_q1_sf100_launches.csv python bench.py --sf 100 --steps 2 --warmup 1 --e2e-steps 0 --cpu-rows 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:sq_agg_small -s 1 -c 1 -f -o gpurun_out/${R}_q1_small python bench.py --sf 100 --steps 2 --warmup 1 --e2e-steps 0 --cpu-rows 0 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_q3_sf10_launches.csv python scripts/q3_time.py 10 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:sq_joinagg -s 1 -c 1 -f -o gpurun_out/${R}_q3_joinagg python scripts/q3_time.py 10 2 > /dev/null 2>&1
SQLRS_B200_TRACE=1 python scripts/q3_time.py 10 3 2>&1 | grep -E "trace|SF" | tail -22 > gpurun_out/${R}_q3_trace.txt
ls -la gpurun_out | tail -20
cat gpurun_out/${R}_bench_n1_sf100.json | cut -c1-400; cat gpurun_out/${R}_bench_q3_sf10.json | cut -c1-600; tail -16 gpurun_out/${R}_q3_trace.txt
