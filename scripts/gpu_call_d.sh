#!/bin/bash
set -x
R=${1:-r01g}
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sq_joinagg|sq_joinprobe|k_join_insert|k_join_probe_emit" -s 4 -c 4 -f -o gpurun_out/${R}_q3_kernels python scripts/q3_time.py 10 3 > /dev/null 2>&1
ls -la gpurun_out/${R}_q3_kernels.ncu-rep
