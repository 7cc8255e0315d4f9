#!/bin/bash
# Tuning sweep of the Q3' join kernels on ONE B200 (run under gpurun): the whole query (scripts/q3_time.py, best of 4 warm runs) under
# different compile-time shapes of the join kernels (SQLRS_B200_JIT_DEFINES, csrc/jit.cpp), Bloom filter sizes and chain-kernel unrolls.
out=${OUT:-gpurun_out/r02h_q3_knobs4.txt}
: > $out
run() {  # run <sf> <defines> <bloom shift> <chain unroll>
  echo "== SF$1 SQLRS_B200_JIT_DEFINES='$2' SQLRS_B200_BLOOM_SHIFT=${3:-1} SQLRS_B200_CUNROLL=${4:-8}" >> $out
  SQLRS_B200_JIT_DEFINES="$2" SQLRS_B200_BLOOM_SHIFT=${3:-1} SQLRS_B200_CUNROLL=${4:-8} timeout 300 python scripts/q3_time.py $1 5 full 2>&1 | grep -E "best|Error|error" | cut -c1-220 >> $out
}
run 100 ""
run 100 "" 1 4
run 100 "" 1 6
run 100 "" 1 12
run 100 "SQ_CMINB=3" 1 12
run 100 "SQ_JUNROLL=6"
run 100 "SQ_JUNROLL=10"
run 100 "SQ_JUNROLL=6;SQ_JMINB=5"
run 10 ""
run 10 "" 1 4
run 10 "SQ_JUNROLL=6"
cat $out
