// sqlrs_b200 JIT skeleton "joinagg": fused  scan -> Filter -> hash-join PROBE -> GROUP BY / aggregate.
// Replaces, for an inner HashJoin directly below an aggregate, the reference's probe loop + build_batch gathers +
// a second build_batch (src/executor/join/hash_join.rs:208-292) and the HashAgg pass over the materialised join
// output (aggregate/hash_agg.rs:33-150): the probe side is scanned ONCE, matches are looked up in the build-side
// table (kernels_join.cu layout, resident in HBM / L2), build-side columns are gathered only for matching rows, and
// every joined row goes straight into the group table — no (build row, probe row) pairs, no joined batch.
//
// Generated in front of this file (after agg_table.cuh):
//   SQ_NKEYS / SQ_NACC / SQ_MATCH_KEYS (aggregate), SQ_JKEYS (join keys), SQ_JMATCH (compare join key tuples),
//   struct SqIn (probe side), struct SqInB (build side),
//   struct SqProbe {pass, h, kb[SQ_JKEYS], knull}; sq_probe_row(in, r, p, e0, e1)   — fused probe-side Filter + join keys
//   struct SqRow {pass, h, kb[K], knull, args...};  sq_row(in, inb, r, b, o, e1)   — join filter + group keys + arguments
//   sq_acc_identity / sq_acc_update / sq_acc_merge_global as in agg.cuh
// HBM-bound on the probe-side scan: algorithmic bytes = 8 B x referenced probe columns per probe row (+ the build
// side once); the table probes are random 8-byte reads that mostly hit L2.

struct SqJoin {           // mirrors sq::JoinTableView (kernels_aot.hpp)
  const i64* slot_rep;    // representative build row per slot, -1 = empty
  const u32* slot_count;
  const u64* slot_start;
  const i64* rows;        // build row ids grouped by slot, ascending
  u32 capacity;
  const u64* h;           // build-side row hashes
  const u64* keys;        // [SQ_JKEYS][n_build] raw key bits (SQ_JMATCH)
  const u32* knull;
  i64 n_build;
  int n_keys;
  int match_keys;
  const u32* build_keep;
};

__device__ __forceinline__ int sq_join_find(const SqJoin& t, const SqProbe& p) {
#if SQ_JMATCH
  if (p.knull != 0u) return -1;  // SQL semantics: a NULL key never joins
#endif
  const u32 mask = t.capacity - 1;
  u32 s = sq_mix32(p.h) & mask;
  for (u32 probes = 0; probes <= mask; probes++) {
    const i64 rep = __ldg(&t.slot_rep[s]);
    if (rep < 0) return -1;
    if (__ldg(&t.h[rep]) == p.h) {
#if SQ_JMATCH
      bool same = true;
#pragma unroll
      for (int k = 0; k < SQ_JKEYS; k++) same = same && (__ldg(&t.keys[(size_t)k * t.n_build + rep]) == p.kb[k]);
      if (same) return (int)s;
#else
      return (int)s;
#endif
    }
    s = (s + 1) & mask;
  }
  return -1;
}

extern "C" __global__ void __launch_bounds__(256) sq_joinagg_kernel(SqIn in, SqInB inb, i64 n, i64 row_base, SqJoin jt, SqTable table, i64 batch_no,
                                                                     u32* __restrict__ status, u32* __restrict__ err) {
  bool any_err = false;
  const int lane = threadIdx.x & 31;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  // warp-uniform trips, two rows per lane per trip so that the probe-side loads of both are in flight together
  for (i64 base = ((i64)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * 2; base < n; base += stride * 2) {
    SqProbe p[2];
    i64 r[2];
    bool live[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      r[u] = base + u * 32 + lane;
      const bool inb_row = r[u] < n;
      bool e0 = false, e1 = false;
      sq_probe_row(in, inb_row ? r[u] : n - 1, p[u], e0, e1);
      live[u] = inb_row && p[u].pass;
      any_err |= (inb_row && e0) || (live[u] && e1);
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      int slot = -1;
      if (live[u]) slot = sq_join_find(jt, p[u]);
      __syncwarp();
      if (slot >= 0) {
        const u32 cnt = __ldg(&jt.slot_count[slot]);
        const i64* brow = jt.rows + __ldg(&jt.slot_start[slot]);
        for (u32 j = 0; j < cnt; j++) {  // per probe row: build rows in insertion order (hash_join.rs:225-235)
          const i64 b = __ldg(&brow[j]);
          SqRow o;
          bool e1 = false;
          sq_row(in, inb, r[u], b, o, e1);
          any_err |= o.pass && e1;
          if (!o.pass) continue;  // non-equi join filter (apply_join_filter, :47-71)
          const int g = sq_table_upsert(table, o.h, o.kb, o.knull);
          if (g < 0) {
            atomicOr(status, SQ_STATUS_FULL);
            continue;
          }
          // first-appearance order of the joined stream = (probe row, match ordinal)
          const u64 ord = ((u64)(row_base + r[u]) << 20) | (u64)(j < 0xfffffu ? j : 0xfffffu);
          if (ord < table.min_row[g]) atomicMin(&table.min_row[g], ord);
          u64 local[SQ_NACC > 0 ? SQ_NACC : 1];
#pragma unroll
          for (int w = 0; w < SQ_NACC; w++) local[w] = sq_acc_identity(w);
          sq_acc_update(local, 1, o);
#pragma unroll
          for (int w = 0; w < SQ_NACC; w++) sq_acc_merge_global(&table.acc[(size_t)w * table.capacity + g], w, local[w], batch_no);
        }
      }
      __syncwarp();
    }
  }
  if (any_err) atomicOr(err, 1u);
}
