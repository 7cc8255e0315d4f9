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

// continue a probe sequence at slot `s` (the slow path behind the batched first probe below)
__device__ __forceinline__ int sq_join_find_from(const SqJoin& t, const SqProbe& p, u32 s) {
  const u32 mask = t.capacity - 1;
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

#ifndef SQ_JUNROLL
#define SQ_JUNROLL 4
#endif

extern "C" __global__ void __launch_bounds__(256) sq_joinagg_kernel(SqIn in, SqInB inb, i64 n, i64 row_base, SqJoin jt, SqTable table, i64 batch_no,
                                                                     u32* __restrict__ status, u32* __restrict__ err) {
  bool any_err = false;
  const int lane = threadIdx.x & 31;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  const u32 jmask = jt.capacity - 1;
  // warp-uniform trips, SQ_JUNROLL rows per lane per trip.  The probe is a chain of dependent random reads
  // (slot -> representative row -> its hash / key), so the first probe of all SQ_JUNROLL rows is issued level by
  // level: SQ_JUNROLL independent reads in flight per thread instead of one chain at a time.
  for (i64 base = ((i64)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * SQ_JUNROLL; base < n; base += stride * SQ_JUNROLL) {
    SqProbe p[SQ_JUNROLL];
    i64 r[SQ_JUNROLL];
    bool live[SQ_JUNROLL];
#pragma unroll
    for (int u = 0; u < SQ_JUNROLL; u++) {
      r[u] = base + u * 32 + lane;
      const bool inb_row = r[u] < n;
      bool e0 = false, e1 = false;
      sq_probe_row(in, inb_row ? r[u] : n - 1, p[u], e0, e1);
      live[u] = inb_row && p[u].pass;
#if SQ_JMATCH
      live[u] = live[u] && p[u].knull == 0u;  // SQL semantics: a NULL key never joins
#endif
      any_err |= (inb_row && e0) || (inb_row && p[u].pass && e1);
    }
    u32 s0[SQ_JUNROLL];
    i64 rep[SQ_JUNROLL];
    u64 hh[SQ_JUNROLL];
#pragma unroll
    for (int u = 0; u < SQ_JUNROLL; u++) {
      s0[u] = sq_mix32(p[u].h) & jmask;
      rep[u] = live[u] ? __ldg(&jt.slot_rep[s0[u]]) : -1;
    }
#pragma unroll
    for (int u = 0; u < SQ_JUNROLL; u++) hh[u] = rep[u] >= 0 ? __ldg(&jt.h[rep[u]]) : 0ULL;
#if SQ_JMATCH
    u64 k0[SQ_JUNROLL];
#pragma unroll
    for (int u = 0; u < SQ_JUNROLL; u++) k0[u] = rep[u] >= 0 ? __ldg(&jt.keys[rep[u]]) : 0ULL;
#endif
#pragma unroll
    for (int u = 0; u < SQ_JUNROLL; u++) {
      int slot = -1;
      if (rep[u] >= 0) {
        bool hit = hh[u] == p[u].h;
#if SQ_JMATCH
        hit = hit && k0[u] == p[u].kb[0];
#pragma unroll
        for (int k = 1; k < SQ_JKEYS; k++) hit = hit && (__ldg(&jt.keys[(size_t)k * jt.n_build + rep[u]]) == p[u].kb[k]);
#endif
        slot = hit ? (int)s0[u] : sq_join_find_from(jt, p[u], (s0[u] + 1) & jmask);
      }
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
