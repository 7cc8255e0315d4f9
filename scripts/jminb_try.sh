#!/bin/bash
# sq_joinagg_kernel at 64 registers (SQ_JMINB=4, the new default) against 80 (SQ_JMINB=1) and 3 CTAs per SM, whole Q3' query; then the
# ncu --set full captures of the Q3' kernels as built now (-> profiles/r02_traffic.json via scripts/make_traffic.py q3), then the Q3' parity tests.
out=gpurun_out/r02h_q3_knobs6.txt
: > $out
for sf in 100 10; do
  for defs in "SQ_JMINB=1" "" "SQ_JMINB=3"; do
    echo "== SF$sf SQLRS_B200_JIT_DEFINES='$defs'" >> $out
    SQLRS_B200_JIT_DEFINES="$defs" timeout 120 python scripts/q3_time.py $sf 5 full 2>&1 | grep -E "best|Error|error" | cut -c1-220 >> $out
  done
done
cat $out
ncu --set full --clock-control none --import-source on -k regex:"sq_joinchain_kernel|sq_joinagg_kernel|sq_joinbuild_kernel" -s 6 -c 3 -o gpurun_out/r02_q3_sf100_kernels -f python scripts/q3_time.py 100 4 full > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sq_joinchain_kernel|sq_joinagg_kernel|sq_joinbuild_kernel" -s 6 -c 3 -o gpurun_out/r02_q3_sf10_kernels -f python scripts/q3_time.py 10 4 full > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "q3" 2>&1 | tail -3
