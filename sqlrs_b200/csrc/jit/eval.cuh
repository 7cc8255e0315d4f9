// sqlrs_b200 JIT skeleton "eval": evaluates the row program for every row and stores its outputs
// as columns — BoundExpr::eval_column (reference src/executor/evaluator.rs:13-28), the Filter
// keep-mask (filter.rs:18-23), join key hashes.  One warp owns 32 consecutive rows, so Boolean
// values and validity bits leave as whole u32 words via __ballot_sync.
//
// Generated in front of this file: SQ_NCOLS, SQ_NOUT, struct SqIn, struct SqOut, struct SqRow,
//   sq_row(in, r, o, e0, e1), sq_store(out, r, inb, lane, o).
// HBM-bound: reads each referenced input column once (8 B/row/column), writes each output once.
#ifndef SQ_EVAL_UNROLL
#define SQ_EVAL_UNROLL 4
#endif

extern "C" __global__ void __launch_bounds__(256) sq_eval_kernel(SqIn in, SqOut out, i64 n, u32* __restrict__ err) {
  const int lane = threadIdx.x & 31;
  const i64 warp = (i64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const i64 nwarps = (i64)gridDim.x * (blockDim.x >> 5);
  bool any_err = false;
  // each warp takes SQ_EVAL_UNROLL groups of 32 rows per trip: the loads of all groups are
  // issued before the first store, which keeps >= SQ_EVAL_UNROLL x columns requests in flight
  for (i64 base = warp * 32 * SQ_EVAL_UNROLL; base < n; base += nwarps * 32 * SQ_EVAL_UNROLL) {
    SqRow o[SQ_EVAL_UNROLL];
    bool inb[SQ_EVAL_UNROLL];
#pragma unroll
    for (int u = 0; u < SQ_EVAL_UNROLL; u++) {
      const i64 r = base + u * 32 + lane;
      inb[u] = r < n;
      bool e0 = false, e1 = false;
      sq_row(in, inb[u] ? r : n - 1, o[u], e0, e1);
      any_err |= inb[u] && (e0 || e1);
    }
#pragma unroll
    for (int u = 0; u < SQ_EVAL_UNROLL; u++) {
      const i64 r = base + u * 32 + lane;
      if (base + u * 32 < n) sq_store(out, r, inb[u], lane, o[u]);  // warp-uniform guard
    }
  }
  if (any_err) atomicOr(err, 1u);
}
