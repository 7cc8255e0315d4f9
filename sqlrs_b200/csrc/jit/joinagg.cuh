// sqlrs_b200 JIT skeleton "joinagg": fused  scan -> Filter -> hash-join PROBE -> GROUP BY / aggregate.
// Replaces, for an inner HashJoin directly below an aggregate, the reference's probe loop + build_batch gathers +
// a second build_batch (src/executor/join/hash_join.rs:208-292) and the HashAgg pass over the materialised join
// output (aggregate/hash_agg.rs:33-150): the probe side is scanned ONCE, matches are looked up in the build-side
// table (kernels_join.cu layout, resident in HBM / L2), build-side columns are gathered only for matching rows, and
// every joined row goes straight into the group table — no (build row, probe row) pairs, no joined batch.
//
// Generated in front of this file (after agg_table.cuh; join_table.cuh follows the generated part):
//   SQ_NKEYS / SQ_NACC / SQ_MATCH_KEYS (aggregate), SQ_JKEYS (join keys), SQ_JMATCH (compare join key tuples),
//   struct SqIn (probe side), struct SqInB (build side),
//   struct SqProbe {pass, h, kb[SQ_JKEYS], knull}; sq_probe_row(in, r, p, e0, e1)   — fused probe-side Filter + join keys
//   struct SqRow {pass, h, kb[K], knull, args...};  sq_row(in, inb, r, b, o, e1)   — join filter + group keys + arguments
//   sq_acc_identity / sq_acc_update / sq_acc_merge_global as in agg.cuh
// HBM-bound on the probe-side scan: algorithmic bytes = 8 B x referenced probe columns per probe row (+ the build
// side once).
//
// Two phases per warp trip (SQ_JUNROLL x 32 rows):
//   A. streaming: every lane evaluates the fused Filter + join-key hash of SQ_JUNROLL rows (coalesced 8-byte loads,
//      all issued before the first use) and tests the hash against the build side's blocked Bloom filter — ONE
//      8-byte read of an L2-resident table per surviving row.  Rows that may match are appended to a per-warp
//      queue in shared memory with ballot/popc (warp-aggregated, no atomics).
//   B. compacted: whenever the queue holds >= 32 candidates a FULL warp probes the hash table (dependent random
//      reads: slot -> representative row -> hash/key), walks the CSR match list, gathers the build columns and
//      upserts into the group table.  So the divergent, latency-bound part runs with all lanes busy however
//      selective the join is, and phase A never waits on it.

#ifndef SQ_JUNROLL
#define SQ_JUNROLL 8
#endif
#define SQ_JBLOCK 256
#ifndef SQ_PREFETCH
#define SQ_PREFETCH 0  // L2 prefetch distance in warp trips (0 = off)
#endif
#ifndef SQ_JMINB
#define SQ_JMINB 4  // __launch_bounds__ minimum CTAs per SM: 64 registers, 32 resident warps (80 registers / 24 warps unconstrained)
#endif
#define SQ_JQUEUE (SQ_JUNROLL * 32 + 32)

// phase B for one candidate row: exact probe, matches in build insertion order, aggregate
__device__ __forceinline__ void sq_joinagg_candidate(const SqIn& in, const SqInB& inb, i64 r, u64 qv, i64 row_base, const SqJoin& jt,
                                                     const SqTable& table, i64 batch_no, u32* status, bool& any_err) {
  SqProbe p;
#if SQ_PQMODE
  sq_probe_unq(qv, p);  // the queued key bits / hash: no second read of the probe row
#else
  bool e0 = false, e1 = false;
  sq_probe_row(in, r, p, e0, e1);  // several compared keys: re-evaluated (the scan's lines are still in L2)
#endif
  i64 rep = -1;
  const int slot = sq_join_find_rep(jt, p, rep);
  if (slot < 0) return;
  // unique build keys (a primary-key side): the slot's representative row IS its match list — three dependent
  // random reads (count, range start, row id) less per match
  const u32 cnt = jt.unique ? 1u : __ldg(&jt.slot_count[slot]);
  const i64* brow = jt.unique ? nullptr : jt.rows + __ldg(&jt.slot_start[slot]);
  for (u32 j = 0; j < cnt; j++) {  // per probe row: build rows in insertion order (hash_join.rs:225-235)
    const i64 b = jt.unique ? rep : __ldg(&brow[j]);
    SqRow o;
    bool e2 = false;
    sq_row(in, inb, r, b, o, e2);
    any_err |= o.pass && e2;
    if (!o.pass) continue;  // non-equi join filter (apply_join_filter, :47-71)
    const int g = sq_table_upsert(table, o.h, o.kb, o.knull);
    if (g < 0) {
      atomicOr(status, SQ_STATUS_FULL);
      continue;
    }
    // first-appearance order of the joined stream = (probe row, match ordinal)
    const u64 ord = ((u64)(row_base + r) << 20) | (u64)(j < 0xfffffu ? j : 0xfffffu);
    if (ord < table.min_row[g]) atomicMin(&table.min_row[g], ord);
    u64 local[SQ_NACC > 0 ? SQ_NACC : 1];
#pragma unroll
    for (int w = 0; w < SQ_NACC; w++) local[w] = sq_acc_identity(w);
    sq_acc_update(local, 1, o);
#pragma unroll
    for (int w = 0; w < SQ_NACC; w++) sq_acc_merge_global(&table.acc[(size_t)w * table.capacity + g], w, local[w], batch_no);
  }
}

extern "C" __global__ void __launch_bounds__(SQ_JBLOCK, SQ_JMINB) sq_joinagg_kernel(SqIn in, SqInB inb, i64 n, i64 row_base, SqJoin jt, SqTable table, i64 batch_no,
                                                                           u32* __restrict__ status, u32* __restrict__ err) {
  __shared__ u32 queue_s[SQ_JBLOCK / 32][SQ_JQUEUE];
  __shared__ u64 queue_vs[SQ_JBLOCK / 32][SQ_PQMODE ? SQ_JQUEUE : 1];
  bool any_err = false;
  const int lane = threadIdx.x & 31;
  u32* queue = queue_s[threadIdx.x >> 5];
  u64* queue_v = queue_vs[threadIdx.x >> 5];
  const u32 lanes_below = (1u << lane) - 1u;
  u32 queued = 0;  // warp-uniform
  const u64 pol_keep = sq_l2_evict_last();
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 base = ((i64)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * SQ_JUNROLL; base < n; base += stride * SQ_JUNROLL) {
#if SQ_PREFETCH
    sq_probe_prefetch(in, base + SQ_PREFETCH * stride * SQ_JUNROLL, n, SQ_JUNROLL * 32, lane);  // this warp's rows SQ_PREFETCH trips ahead
#endif
    SQ_PHASE_A(SQ_JUNROLL, base + u * 32 + lane)  // n < 2^32 (checked by the host)
    // ---- phase B: full warps only
    while (queued >= 32) {
      queued -= 32;
      sq_joinagg_candidate(in, inb, (i64)queue[queued + lane], SQ_PQMODE ? queue_v[queued + lane] : 0ULL, row_base, jt, table, batch_no, status, any_err);
      __syncwarp();
    }
  }
  if ((u32)lane < queued) sq_joinagg_candidate(in, inb, (i64)queue[lane], SQ_PQMODE ? queue_v[lane] : 0ULL, row_base, jt, table, batch_no, status, any_err);
  if (any_err) atomicOr(err, 1u);
}

#if SQ_TMA
// ---------------------------------------------------------------------------------------------------------------
// TMA variant (opt-in: SQLRS_B200_TMA=1): "bulk-async staging of probe-side batches into shared memory", warp-specialised.
//   * ONE producer warp per CTA: its lane 0 issues cp.async.bulk copies of SQ_TROWS-row slices of the columns the probe
//     program reads (Filter column(s) + join key(s), 8 KB per column) into a SQ_TSTAGES-deep shared-memory ring.  Per stage
//     a FULL mbarrier (armed with the byte count, completed by the copies) and an EMPTY mbarrier (one arrival per consumer
//     warp).  The producer only ever waits for a stage to be EMPTY, so up to SQ_TSTAGES tiles are in flight per CTA with no
//     registers spent on loads.
//   * SQ_TCONSUMERS consumer warps: each waits for FULL, evaluates its 1/SQ_TCONSUMERS slice of the tile out of shared
//     memory (phase A), and releases the stage with ONE mbarrier.arrive per warp as soon as its values are in registers —
//     no __syncthreads(), nobody waits for another warp.  The Bloom test and phase B (exact probe of the queued candidates,
//     gathers, group upsert) run AFTER the release and touch global memory only, as in the register-staged kernel.
// Round 1's variant (two stages, issuing thread inside a consumer warp, a CTA barrier per tile) coupled every warp to the
// slowest phase-B warp; this is the structure the comparison in profiles/r02g_* is made with.
// Requirements checked by the host (ops_agg.cpp): every staged column is 8 bytes wide and 16-byte aligned; the rows behind
// the last full tile go through sq_joinagg_kernel.
#ifndef SQ_TSTAGES
#define SQ_TSTAGES 4
#endif
#ifndef SQ_TCONSUMERS
#define SQ_TCONSUMERS 8
#endif
#define SQ_TBLOCK (32 * (SQ_TCONSUMERS + 1))
#define SQ_TUNROLL (SQ_TROWS / SQ_TCONSUMERS / 32)
#define SQ_TQUEUE (SQ_TUNROLL * 32 + 32)
__device__ __forceinline__ u32 sq_smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sq_mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sq_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sq_mbar_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sq_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sq_mbar_arrive(u64* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sq_smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void sq_bulk_g2s(void* dst, const void* src, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sq_smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(sq_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void sq_mbar_wait(u64* bar, u32 parity) {
  u32 done = 0;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(sq_smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!done);
}

extern "C" __global__ void __launch_bounds__(SQ_TBLOCK) sq_joinagg_tma_kernel(SqIn in, SqInB inb, i64 n_tiles, i64 row_base, SqJoin jt, SqTable table,
                                                                               i64 batch_no, u32* __restrict__ status, u32* __restrict__ err) {
  extern __shared__ __align__(128) u64 sq_tiles[];  // [SQ_TSTAGES][SQ_TILE_NCOLS][SQ_TROWS]
  __shared__ __align__(8) u64 bar_full[SQ_TSTAGES], bar_empty[SQ_TSTAGES];
  __shared__ u32 queue_s[SQ_TCONSUMERS][SQ_TQUEUE];
  __shared__ u64 queue_vs[SQ_TCONSUMERS][SQ_PQMODE ? SQ_TQUEUE : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < SQ_TSTAGES; s++) {
      sq_mbar_init(&bar_full[s], 1);
      sq_mbar_init(&bar_empty[s], SQ_TCONSUMERS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == SQ_TCONSUMERS) {
    // ---- producer warp
    if (lane == 0) {
      const int tile_cols[SQ_TILE_NCOLS] = SQ_TILE_COLS;
      u32 k = 0;
      for (i64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, k++) {
        const u32 stage = k % SQ_TSTAGES, use = k / SQ_TSTAGES;
        sq_mbar_wait(&bar_empty[stage], (use & 1u) ^ 1u);  // a fresh barrier passes: the stage has never been filled
        sq_mbar_expect_tx(&bar_full[stage], SQ_TILE_NCOLS * SQ_TROWS * 8);
#pragma unroll
        for (int c = 0; c < SQ_TILE_NCOLS; c++)
          sq_bulk_g2s(sq_tiles + ((size_t)stage * SQ_TILE_NCOLS + c) * SQ_TROWS, (const u64*)in.col[tile_cols[c]] + tile * SQ_TROWS, SQ_TROWS * 8, &bar_full[stage]);
      }
    }
    return;
  }
  // ---- consumer warps
  bool any_err = false;
  u32* queue = queue_s[warp];
  u64* queue_v = queue_vs[warp];
  const u32 lanes_below = (1u << lane) - 1u;
  const u64 pol_keep = sq_l2_evict_last();
  u32 queued = 0;  // warp-uniform
  u32 k = 0;
  for (i64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, k++) {
    const u32 stage = k % SQ_TSTAGES, use = k / SQ_TSTAGES;
    sq_mbar_wait(&bar_full[stage], use & 1u);
    const u64* tl = sq_tiles + (size_t)stage * SQ_TILE_NCOLS * SQ_TROWS;
    const int t0 = warp * (SQ_TUNROLL * 32);
    const i64 base = tile * SQ_TROWS + t0;
    // ---- phase A out of shared memory
    u64 qv[SQ_TUNROLL];
    u32 bits[SQ_TUNROLL], bw[SQ_TUNROLL];
    bool live[SQ_TUNROLL];
#pragma unroll
    for (int u = 0; u < SQ_TUNROLL; u++) {
      SqProbe p;
      bool e0 = false, e1 = false;
      sq_probe_row_tile(in, tl, t0 + u * 32 + lane, base + u * 32 + lane, p, e0, e1);
      live[u] = p.pass;
#if SQ_JMATCH
      live[u] = live[u] && p.knull == 0u;
#endif
      any_err |= e0 || (p.pass && e1);
      qv[u] = SQ_PQMODE ? sq_probe_qv(p) : 0ULL;
      bits[u] = sq_bloom_bits(p.h);
      bw[u] = sq_bloom_word(p.h, jt.bloom_mask);
    }
    __syncwarp();
    if (lane == 0) sq_mbar_arrive(&bar_empty[stage]);  // this warp is done with the stage: the producer may refill it
#pragma unroll
    for (int u = 0; u < SQ_TUNROLL; u++) bw[u] = live[u] ? sq_ld_u32_l2(&jt.bloom[bw[u]], pol_keep) : 0u;
#pragma unroll
    for (int u = 0; u < SQ_TUNROLL; u++) {
      const bool cand = live[u] && (bw[u] & bits[u]) == bits[u];
      const u32 m = __ballot_sync(0xffffffffu, cand);
      if (cand) {
        const u32 pos = queued + __popc(m & lanes_below);
        queue[pos] = (u32)(base + u * 32 + lane);
        if (SQ_PQMODE) queue_v[pos] = qv[u];
      }
      queued += __popc(m);
    }
    __syncwarp();
    // ---- phase B: full warps only
    while (queued >= 32) {
      queued -= 32;
      sq_joinagg_candidate(in, inb, (i64)queue[queued + lane], SQ_PQMODE ? queue_v[queued + lane] : 0ULL, row_base, jt, table, batch_no, status, any_err);
      __syncwarp();
    }
  }
  if ((u32)lane < queued) sq_joinagg_candidate(in, inb, (i64)queue[lane], SQ_PQMODE ? queue_v[lane] : 0ULL, row_base, jt, table, batch_no, status, any_err);
  if (any_err) atomicOr(err, 1u);
}
#endif  // SQ_TMA
