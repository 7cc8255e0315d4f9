#!/bin/bash
set -x
R=${1:-r01e}
SF=${2:-100}
mkdir -p gpurun_out
SQLRS_B200_TRACE=1 timeout 300 python scripts/q3_time.py $SF 4 2>&1 | grep -E "trace|SF" | tail -64 > gpurun_out/${R}_q3_sf${SF}_trace.txt; cat gpurun_out/${R}_q3_sf${SF}_trace.txt
