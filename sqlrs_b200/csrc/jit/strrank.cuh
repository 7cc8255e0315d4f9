// sqlrs_b200 JIT building block "strrank" (placed behind the prelude of a kernel whose row program compares Utf8 values by
// order, csrc/jit.cpp): Utf8 values are string-pool ids (csrc/device.hpp); `<, <=, >, >=` go through the pool's byte-wise rank
// table — Rust's `str` order, what arrow's lt_utf8 / gt_utf8 ... compare by (reference: gt_dyn / lt_dyn / gt_eq_dyn / lt_eq_dyn,
// src/executor/array_compute.rs:80-83).  The host points sq_rank_table at the current table before every launch (jit_launch).
__device__ const int* sq_rank_table;
__device__ __forceinline__ int sq_str_rank(i64 id) { return __ldg(sq_rank_table + id); }
