#!/bin/bash
# Tuning sweep of the Q3' join kernels on ONE B200 (run under gpurun): the whole query (scripts/q3_time.py, best of 4 warm runs) under
# different compile-time shapes of the join kernels (SQLRS_B200_JIT_DEFINES, csrc/jit.cpp), Bloom filter sizes, chain-kernel unrolls
# and load factors of the key-in-slot tables.
out=${OUT:-gpurun_out/r02h_q3_knobs5.txt}
: > $out
run() {  # run <sf> <defines> <bloom shift> <chain unroll> <kv max load %>
  echo "== SF$1 SQLRS_B200_JIT_DEFINES='$2' SQLRS_B200_BLOOM_SHIFT=${3:-1} SQLRS_B200_CUNROLL=${4:-8} SQLRS_B200_KV_MAXLOAD=${5:-50}" >> $out
  SQLRS_B200_JIT_DEFINES="$2" SQLRS_B200_BLOOM_SHIFT=${3:-1} SQLRS_B200_CUNROLL=${4:-8} SQLRS_B200_KV_MAXLOAD=${5:-50} timeout 300 python scripts/q3_time.py $1 5 full 2>&1 | grep -E "best|Error|error" | cut -c1-220 >> $out
}
run 100 "" 1 8 50
run 100 "" 1 8 70
run 100 "" 1 8 85
run 100 "" 1 8 95
run 10 "" 1 8 50
run 10 "" 1 8 85
cat $out
