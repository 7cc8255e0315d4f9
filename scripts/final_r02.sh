#!/bin/bash
# Round-2 evidence on ONE B200 (run under gpurun): the GPU test-suite log, the default bench line, the ncu launch list of the same
# bench command, and one ncu --set full capture of the dominant kernels of Q1' SF100 and Q3' SF10 / SF100 (-> profiles/r02_traffic.json).
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/r02_pytest_gpu.txt
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --e2e-steps 0 --cpu-rows 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sq_agg_small" -s 4 -c 1 -o gpurun_out/r02_q1_sf100_sq_agg_small python bench.py --steps 2 --warmup 3 --e2e-steps 0 --cpu-rows 0 --q3 off > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sq_joinchain_kernel|sq_joinagg_kernel|sq_joinbuild_kernel" -s 6 -c 3 -o gpurun_out/r02_q3_sf100_kernels python scripts/q3_time.py 100 4 full > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sq_joinchain_kernel|sq_joinagg_kernel|sq_joinbuild_kernel" -s 6 -c 3 -o gpurun_out/r02_q3_sf10_kernels python scripts/q3_time.py 10 4 full > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
