#!/bin/bash
# Tuning sweep of the Q3' join kernels on ONE B200 (run under gpurun): the whole query (scripts/q3_time.py, best of 4 warm runs) under
# different compile-time shapes of the join kernels (SQLRS_B200_JIT_DEFINES, csrc/jit.cpp) and Bloom filter sizes.
out=gpurun_out/r02h_q3_knobs3.txt
: > $out
run() {  # run <sf> <defines> <bloom shift>
  echo "== SF$1 SQLRS_B200_JIT_DEFINES='$2' SQLRS_B200_BLOOM_SHIFT=${3:-1}" >> $out
  SQLRS_B200_JIT_DEFINES="$2" SQLRS_B200_BLOOM_SHIFT=${3:-1} timeout 300 python scripts/q3_time.py $1 5 full 2>&1 | grep -E "best|Error|error" | cut -c1-220 >> $out
}
run 100 ""
run 100 "SQ_PREFETCH=1"
run 100 "SQ_PREFETCH=2"
run 100 "SQ_PREFETCH=4"
run 10 ""
run 10 "SQ_PREFETCH=1"
run 10 "SQ_PREFETCH=2"
cat $out
