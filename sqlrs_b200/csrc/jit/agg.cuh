// sqlrs_b200 JIT skeleton "agg": fused Filter -> expression -> GROUP BY / aggregate over one batch.
// Replaces the reference's per-batch "hash rows -> HashMap -> take per group -> update_batch"
// (src/executor/aggregate/hash_agg.rs:33-150, simple_agg.rs:27-65, accumulators sum.rs / count.rs /
// min_max.rs) and, when the child is a Filter, filter.rs:14-26 as well: every input column is read
// once, nothing is materialised.
//
// Generated in front of this file:
//   SQ_NCOLS, SQ_NKEYS (K), SQ_NACC (W accumulator words per group), SQ_MATCH_KEYS, SQ_BLOCK (T),
//   SQ_SLOTS (S), SQ_UNROLL (R), struct SqIn, struct SqRow {pass, h, kb[K], knull, args...},
//   sq_row(in, r, o, e0, e1), sq_acc_identity(w), sq_acc_update(a, stride, o),
//   sq_acc_reduce(w, x, y), sq_acc_merge_global(p, w, x, batch_no)
//
// Two kernels over the same row program:
//  * sq_agg_small  — few groups (<= SQ_SLOTS per CTA): every thread owns a PRIVATE copy of the
//    accumulators in shared memory ([word][slot][thread], 8-byte lanes -> conflict-free), so the hot
//    loop has no atomics at all; group slots come from a tiny CTA-shared open-addressed table.
//    CTA partials go to a scratch area, sq_agg_merge folds them into the operator's table in HBM.
//    HBM-bound: algorithmic bytes = 8 B x referenced columns per row.
//  * sq_agg_global — many groups: open-addressed table in HBM, one find-or-insert + native 64-bit
//    atomics per row and accumulator word.
// Group identity = the reference's 64-bit row hash (quirk K2), optionally plus the key tuple
// (SQLRS_MATCH_HASH_AND_KEY).  First-appearance order is kept as the minimum global row id.

// (struct SqTable / SqPartial, status bits and sq_table_upsert come from agg_table.cuh, compiled in front of this file)

// ------------------------------------------------------------------------------------------
// shared-memory layout of sq_agg_small
//   u64 acc[(W+1)][S][T]   private accumulators, word W = min row id
//   CTA-shared slot table with SQ_TSLOTS = 8*S entries (load factor <= 1/8) probed through a branch-free
//   two-slot window: u64 ttag[TS]; u64 thash[TS]; u64 tkeys[TS][K]; u32 tknull[TS]; u32 tstate[TS]; u32 tgroup[TS]
//   u32 ngroups; u32 flags
#define SQ_ACC_WORDS (SQ_NACC + 1)
#define SQ_TSLOTS (8 * SQ_SLOTS)

struct SqSlotTable {
  u64* ttag;    // 0 = not published; else (hash | 1): publication state + 63 hash bits in ONE word for the hot-path window
  u64* thash;
  u64* tkeys;
  u32* tknull;
  u32* tstate;  // 0 empty, 1 being written, 2 ready
  u32* tgroup;  // dense group index < SQ_SLOTS (position of the private accumulators)
  u32* ngroups;
};

// ---- CTA-shared slot table, two-level protocol.
// Hot path: sq_small_find — read-only probe, no waiting: the dense group index, or -1 when the key is not
// (yet) published.  Cold path (each CTA meets each group once): sq_small_insert — blocking find-or-insert run by
// ONE elected lane per warp at a time (see the warp-uniform loop in the kernel), so a lane never spins on a slot
// that a sibling lane of its own warp is still writing; the writer it may wait for is always in another warp.
struct SqKey {  // identity of a row's group, passed BY VALUE to the out-of-line cold path
  u64 h;
  u64 kb[SQ_NKEYS > 0 ? SQ_NKEYS : 1];
  u32 knull;
};
template <typename R>
__device__ __forceinline__ bool sq_small_same(const SqSlotTable& t, u32 s, const R& o) {
  if (*((volatile u64*)&t.thash[s]) != o.h) return false;
#if SQ_MATCH_KEYS
  bool same = *((volatile u32*)&t.tknull[s]) == o.knull;
#pragma unroll
  for (int k = 0; k < SQ_NKEYS; k++) same = same && (*((volatile u64*)&t.tkeys[s * (SQ_NKEYS > 0 ? SQ_NKEYS : 1) + k]) == o.kb[k]);
  return same;
#else
  return true;
#endif
}

// The first TWO slots of the probe sequence are examined without a data-dependent branch: one tag word per slot
// carries "published" and 63 bits of the hash, then ONE full comparison at the selected slot — so a key displaced by one
// position costs exactly what a key in its home slot costs.  Measured reason: with one-slot-at-a-time probing the Q1'
// kernel swung between 6.4 and 9.1 ms at SF100 depending on whether the 8 group keys happened to land in 8 distinct home
// slots (a change of the placement hash exposed it) — every divergent extra probe is paid by the whole warp.  Longer
// chains (rare at load <= 1/8) continue in the loop.
template <int TS, typename R>
__device__ __forceinline__ int sq_small_find(const SqSlotTable& t, const R& o) {
#if SQ_NKEYS == 0
  return 0;
#else
  const u32 s0 = sq_mix32(o.h) & (TS - 1), s1 = (s0 + 1) & (TS - 1);
  const u64 want = o.h | 1ULL;
  const u64 t0 = *((volatile u64*)&t.ttag[s0]), t1 = *((volatile u64*)&t.ttag[s1]);
  const bool tag0 = t0 == want, tag1 = t0 != 0ULL && t1 == want;
  const u32 sel = tag0 ? s0 : s1;
  if ((tag0 || tag1) && sq_small_same(t, sel, o)) return (int)*((volatile u32*)&t.tgroup[sel]);
  if (t0 == 0ULL || t1 == 0ULL) return -1;  // the chain ends inside the window: not published (yet)
  if (tag0 && tag1 && sq_small_same(t, s1, o)) return (int)*((volatile u32*)&t.tgroup[s1]);  // equal tags, different keys
  u32 s = (s1 + 1) & (TS - 1);
  for (int probes = 2; probes < TS; probes++) {
    if (*((volatile u32*)&t.tstate[s]) != 2u) return -1;  // empty or being written: not published (yet)
    if (sq_small_same(t, s, o)) return (int)*((volatile u32*)&t.tgroup[s]);
    s = (s + 1) & (TS - 1);
  }
  return -1;
#endif
}

// The same lookup for a CONVERGED warp (every lane calls it, `live` says whose result counts): when every live lane
// finds its key's tag in the home slot — the normal state at load <= 1/8 — a warp-uniform vote skips the second slot's
// loads altogether; otherwise the whole warp takes the two-slot window together.  Either way no lane waits for another
// lane's longer probe sequence.
template <int TS, typename R>
__device__ __forceinline__ int sq_small_find_warp(const SqSlotTable& t, const R& o, bool live) {
#if SQ_NKEYS == 0
  return live ? 0 : -1;
#else
  const u32 s0 = sq_mix32(o.h) & (TS - 1);
  const u64 want = o.h | 1ULL;
  const bool tag0 = *((volatile u64*)&t.ttag[s0]) == want;
  if (__all_sync(SQ_FULL, !live || tag0)) {
    if (live && sq_small_same(t, s0, o)) return (int)*((volatile u32*)&t.tgroup[s0]);
    return live ? sq_small_find<TS>(t, o) : -1;  // equal tag, different key (a 63-bit collision): the general path
  }
  return live ? sq_small_find<TS>(t, o) : -1;
#endif
}

// blocking find-or-insert; -1 = more than SQ_SLOTS groups (the slot is published as unusable)
template <int TS, int NG>
__device__ __forceinline__ int sq_small_insert(const SqSlotTable& t, const SqKey& o) {
#if SQ_NKEYS == 0
  return 0;
#else
  u32 s = sq_mix32(o.h) & (TS - 1);
  for (int probes = 0; probes < TS;) {
    const u32 st = *((volatile u32*)&t.tstate[s]);
    if (st == 2u) {
      if (sq_small_same(t, s, o)) {
        const u32 g = *((volatile u32*)&t.tgroup[s]);
        return g < NG ? (int)g : -1;
      }
      s = (s + 1) & (TS - 1);
      probes++;
      continue;
    }
    if (st == 0u && atomicCAS(&t.tstate[s], 0u, 1u) == 0u) {
      const u32 g = atomicAdd(t.ngroups, 1u);
      t.thash[s] = o.h;
#pragma unroll
      for (int k = 0; k < SQ_NKEYS; k++) t.tkeys[s * SQ_NKEYS + k] = o.kb[k];
      t.tknull[s] = o.knull;
      t.tgroup[s] = g < NG ? g : 0xffffffffu;
      __threadfence_block();
      *((volatile u64*)&t.ttag[s]) = o.h | 1ULL;
      atomicExch(&t.tstate[s], 2u);
      return g < NG ? (int)g : -1;
    }
    // st == 1 (a lane of ANOTHER warp is publishing this slot) or the claim was lost: look again
  }
  return -1;
#endif
}

// group index for every live lane of the warp (call with the whole warp converged)
// cold path of sq_small_group: some lane of the (converged) warp met an unpublished key
template <int TS, int NG>
__device__ __noinline__ int sq_small_resolve(const SqSlotTable t, const SqKey o, int g, bool need) {  // -2 = overflow
  bool overflow = false;
  while (__any_sync(SQ_FULL, need)) {
    const int leader = __ffs(__ballot_sync(SQ_FULL, need)) - 1;
    if ((int)(threadIdx.x & 31) == leader) {
      g = sq_small_insert<TS, NG>(t, o);
      if (g < 0) overflow = true;
      need = false;
    }
    __syncwarp();
    if (need) {  // the leader may just have published this lane's key
      g = sq_small_find<TS>(t, o);
      need = g < 0;
    }
  }
  return overflow ? -2 : g;
}

// group index for every live lane of the warp (call with the whole warp converged)
template <int TS, int NG>
__device__ __forceinline__ int sq_small_group(const SqSlotTable& t, const SqRow& o, bool live, bool& overflow) {
  int g = live ? sq_small_find<TS>(t, o) : 0;
  const bool need = live && g < 0;  // not published yet (or an overflowed slot, which reads back as -1)
  if (__any_sync(SQ_FULL, need)) {
    SqKey key;
    key.h = o.h;
    key.knull = o.knull;
#pragma unroll
    for (int k = 0; k < (SQ_NKEYS > 0 ? SQ_NKEYS : 1); k++) key.kb[k] = o.kb[k];
    g = sq_small_resolve<TS, NG>(t, key, g, need);
    if (g == -2) {
      overflow = true;
      g = -1;
    }
  }
  return live ? g : -1;
}

// One tile processed with the blocking protocol: called (out of line, by the whole converged warp) only when some
// lane met a key that is not published yet — a handful of tiles per CTA.  Re-evaluates the tile's rows so that
// nothing of the hot loop has to stay live across the call.  Returns bit 0 = error flag, bit 1 = overflow.
__device__ __noinline__ u32 sq_small_tile_slow(const SqIn in, i64 n, i64 row_base, i64 base, u64* acc, const SqSlotTable tab) {
  const int tid = threadIdx.x;
  u32 ret = 0;
  for (int u = 0; u < SQ_UNROLL; u++) {
    const i64 r = base + (i64)u * SQ_BLOCK + tid;
    const bool inb = r < n;
    SqRow o;
    bool e0 = false, e1 = false;
    sq_row(in, inb ? r : n - 1, o, e0, e1);
    const bool live = inb && o.pass;
    if ((inb && e0) || (live && e1)) ret |= 1u;
    bool ovf = false;
    const int g = sq_small_group<SQ_TSLOTS, SQ_SLOTS>(tab, o, live, ovf);
    if (ovf) ret |= 2u;
    __syncwarp();
    if (g >= 0) {
      u64* a = acc + (size_t)g * SQ_BLOCK + tid;
      sq_acc_update(a, SQ_SLOTS * SQ_BLOCK, o);
      u64* mr = a + (size_t)SQ_NACC * SQ_SLOTS * SQ_BLOCK;
      const u64 gr = (u64)(row_base + r);
      if (gr < *mr) *mr = gr;
    }
    __syncwarp();
  }
  return ret;
}

extern "C" __global__ void __launch_bounds__(SQ_BLOCK, SQ_MINCTAS) sq_agg_small(SqIn in, i64 n, i64 row_base, SqPartial part,
                                                                     u32* __restrict__ status, u32* __restrict__ err) {
  extern __shared__ __align__(16) unsigned char sq_smem[];
  u64* acc = (u64*)sq_smem;
  SqSlotTable tab;
  tab.ttag = acc + (size_t)SQ_ACC_WORDS * SQ_SLOTS * SQ_BLOCK;
  tab.thash = tab.ttag + SQ_TSLOTS;
  tab.tkeys = tab.thash + SQ_TSLOTS;
  tab.tknull = (u32*)(tab.tkeys + SQ_TSLOTS * (SQ_NKEYS > 0 ? SQ_NKEYS : 1));
  tab.tstate = tab.tknull + SQ_TSLOTS;
  tab.tgroup = tab.tstate + SQ_TSLOTS;
  tab.ngroups = tab.tgroup + SQ_TSLOTS;
  u32* flags = tab.ngroups + 1;
  const int tid = threadIdx.x;

  if ((*((volatile u32*)status) & SQ_STATUS_OVERFLOW) != 0u) return;  // sticky: the global kernel owns this operator now

#pragma unroll
  for (int w = 0; w < SQ_ACC_WORDS; w++) {
    const u64 ident = (w == SQ_NACC) ? SQ_EMPTY_ROW : sq_acc_identity(w);
#pragma unroll
    for (int s = 0; s < SQ_SLOTS; s++) acc[((size_t)w * SQ_SLOTS + s) * SQ_BLOCK + tid] = ident;
  }
  for (int s = tid; s < SQ_TSLOTS; s += SQ_BLOCK) {
    tab.tstate[s] = 0u;
    tab.ttag[s] = 0ULL;
  }
  if (tid == 0) {
    *tab.ngroups = 0u;
    *flags = 0u;
  }
  __syncthreads();

  bool any_err = false;
  bool overflow = false;
  const i64 tile = (i64)SQ_BLOCK * SQ_UNROLL;
  for (i64 base = (i64)blockIdx.x * tile; base < n; base += (i64)gridDim.x * tile) {
    if (*((volatile u32*)flags) != 0u || (*((volatile u32*)status) & SQ_STATUS_OVERFLOW) != 0u) break;  // some CTA overflowed: the batch is re-run anyway
    // all loads of the tile are issued before the first use: SQ_UNROLL x columns requests in flight per thread
    SqRow o[SQ_UNROLL];
    bool live[SQ_UNROLL];
#pragma unroll
    for (int u = 0; u < SQ_UNROLL; u++) {
      const i64 r = base + (i64)u * SQ_BLOCK + tid;
      const bool inb = r < n;
      bool e0 = false, e1 = false;
      sq_row(in, inb ? r : n - 1, o[u], e0, e1);
      live[u] = inb && o[u].pass;
      any_err |= (inb && e0) || (live[u] && e1);
    }
    // group lookup, hot path: read-only probes of the CTA-shared slot table, no waiting, no calls
    int g[SQ_UNROLL];
    bool miss = false;
#pragma unroll
    for (int u = 0; u < SQ_UNROLL; u++) {
      g[u] = sq_small_find_warp<SQ_TSLOTS>(tab, o[u], live[u]);
      miss |= live[u] && g[u] < 0;
    }
    // the probes are the only divergent code: reconverge (without this the warp stays split per group and every
    // later load is replayed per fragment), then decide warp-uniformly
    if (__any_sync(SQ_FULL, miss)) {  // an unpublished key: start-up tiles only
      const u32 ret = sq_small_tile_slow(in, n, row_base, base, acc, tab);
      any_err |= (ret & 1u) != 0u;
      if (ret & 2u) {
        overflow = true;
        *((volatile u32*)flags) = 1u;
        atomicOr(status, SQ_STATUS_OVERFLOW);  // lets the other CTAs stop early
      }
      continue;
    }
#pragma unroll
    for (int u = 0; u < SQ_UNROLL; u++) {
      if (g[u] >= 0) {
        u64* a = acc + (size_t)g[u] * SQ_BLOCK + tid;
        sq_acc_update(a, SQ_SLOTS * SQ_BLOCK, o[u]);
        u64* mr = a + (size_t)SQ_NACC * SQ_SLOTS * SQ_BLOCK;
        const u64 gr = (u64)(row_base + base + (i64)u * SQ_BLOCK + tid);
        if (gr < *mr) *mr = gr;
      }
    }
    __syncwarp();
  }
  if (any_err) atomicOr(err, 1u);
  if (overflow) *((volatile u32*)flags) = 1u;
  __syncthreads();
  if (*((volatile u32*)flags) != 0u) {  // the host reruns this batch on the HBM table
    if (tid == 0) atomicOr(status, SQ_STATUS_OVERFLOW);
    if (tid < SQ_SLOTS) part.state[(size_t)blockIdx.x * SQ_SLOTS + tid] = 0u;
    return;
  }

  // fold the T private copies: one warp per (word, group), lanes stride over the threads
  const int lane = tid & 31, wid = tid >> 5;
  for (int item = wid; item < SQ_ACC_WORDS * SQ_SLOTS; item += SQ_BLOCK / 32) {
    const int w = item / SQ_SLOTS, s = item % SQ_SLOTS;
    const u64* src = acc + ((size_t)w * SQ_SLOTS + s) * SQ_BLOCK;
    u64 x = src[lane];
    for (int t = lane + 32; t < SQ_BLOCK; t += 32) x = (w == SQ_NACC) ? (src[t] < x ? src[t] : x) : sq_acc_reduce(w, x, src[t]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const u64 y = __shfl_xor_sync(SQ_FULL, x, d);
      x = (w == SQ_NACC) ? (y < x ? y : x) : sq_acc_reduce(w, x, y);
    }
    if (lane == 0) {
      const size_t e = (size_t)blockIdx.x * SQ_SLOTS + s;
      if (w == SQ_NACC) part.min_row[e] = x;
      else part.acc[e * SQ_NACC + w] = x;
    }
  }
  // identity of each dense group: scatter from the slot table
  if (tid < SQ_SLOTS) part.state[(size_t)blockIdx.x * SQ_SLOTS + tid] = 0u;
  __syncthreads();
#if SQ_NKEYS > 0
  for (int s = tid; s < SQ_TSLOTS; s += SQ_BLOCK) {
    if (tab.tstate[s] != 2u) continue;
    const u32 g = tab.tgroup[s];
    if (g >= SQ_SLOTS) continue;
    const size_t e = (size_t)blockIdx.x * SQ_SLOTS + g;
    part.state[e] = 2u;
    part.hash[e] = tab.thash[s];
    part.knull[e] = tab.tknull[s];
#pragma unroll
    for (int k = 0; k < SQ_NKEYS; k++) part.keys[e * SQ_NKEYS + k] = tab.tkeys[s * SQ_NKEYS + k];
  }
#else
  if (tid == 0) {
    const size_t e = (size_t)blockIdx.x * SQ_SLOTS;
    part.state[e] = 2u;
    part.hash[e] = 0ULL;
    part.knull[e] = 0u;
  }
#endif
}

// folds the CTA partials of one batch into the operator's table; one thread per (cta, slot)
extern "C" __global__ void __launch_bounds__(128) sq_agg_merge(SqPartial part, int n_entries, SqTable table, i64 batch_no,
                                                                u32* __restrict__ status) {
  if ((*((volatile u32*)status) & (SQ_STATUS_OVERFLOW | SQ_STATUS_OVERFLOW2)) != 0u) return;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_entries) return;
  if (part.state[e] != 2u) return;
  const u64 mr = part.min_row[e];
  if (mr == SQ_EMPTY_ROW) return;  // slot claimed but no surviving row (cannot happen; defensive)
  u64 kb[SQ_NKEYS > 0 ? SQ_NKEYS : 1];
#pragma unroll
  for (int k = 0; k < SQ_NKEYS; k++) kb[k] = part.keys[(size_t)e * SQ_NKEYS + k];
  const int slot = sq_table_upsert(table, part.hash[e], kb, part.knull[e]);
  if (slot < 0) {
    atomicOr(status, SQ_STATUS_FULL);
    return;
  }
  atomicMin(&table.min_row[slot], mr);
#pragma unroll
  for (int w = 0; w < SQ_NACC; w++)
    sq_acc_merge_global(&table.acc[(size_t)w * table.capacity + slot], w, part.acc[(size_t)e * SQ_NACC + w], batch_no);
}

// ------------------------------------------------------------------------------------------
// sq_agg_medium — up to SQ_MSLOTS groups per CTA (hundreds): ONE copy of the accumulators per CTA in shared
// memory, updated with shared-memory atomics (contention is spread over many groups here, unlike the
// few-group case that needs private copies), then flushed into the HBM table once per CTA.
//   u64 macc[(W+1)][M]; slot table with 2*M entries: thash, tkeys[K], tknull, tstate, tgroup; ngroups; flags
#define SQ_MTSLOTS (2 * SQ_MSLOTS)

// the medium kernel's slow tile (see sq_small_tile_slow): blocking protocol + shared-memory atomics
__device__ __noinline__ u32 sq_medium_tile_slow(const SqIn in, i64 n, i64 row_base, i64 base, u64* macc, const SqSlotTable tab) {
  const int tid = threadIdx.x;
  u32 ret = 0;
  for (int u = 0; u < SQ_MUNROLL; u++) {
    const i64 r = base + (i64)u * 256 + tid;
    const bool inb = r < n;
    SqRow o;
    bool e0 = false, e1 = false;
    sq_row(in, inb ? r : n - 1, o, e0, e1);
    const bool live = inb && o.pass;
    if ((inb && e0) || (live && e1)) ret |= 1u;
    bool ovf = false;
    const int g = sq_small_group<SQ_MTSLOTS, SQ_MSLOTS>(tab, o, live, ovf);
    if (ovf) ret |= 2u;
    __syncwarp();
    if (g >= 0) {
      u64 local[SQ_NACC > 0 ? SQ_NACC : 1];
#pragma unroll
      for (int w = 0; w < SQ_NACC; w++) local[w] = sq_acc_identity(w);
      sq_acc_update(local, 1, o);
#pragma unroll
      for (int w = 0; w < SQ_NACC; w++) sq_acc_merge_shared(&macc[(size_t)w * SQ_MSLOTS + g], w, local[w]);
      atomicMin(&macc[(size_t)SQ_NACC * SQ_MSLOTS + g], (u64)(row_base + r));
    }
    __syncwarp();
  }
  return ret;
}

extern "C" __global__ void __launch_bounds__(256) sq_agg_medium(SqIn in, i64 n, i64 row_base, SqPartial part,
                                                                 u32* __restrict__ status, u32* __restrict__ err) {
  extern __shared__ __align__(16) unsigned char sq_smem[];
  u64* macc = (u64*)sq_smem;
  SqSlotTable tab;
  tab.ttag = macc + (size_t)SQ_ACC_WORDS * SQ_MSLOTS;
  tab.thash = tab.ttag + SQ_MTSLOTS;
  tab.tkeys = tab.thash + SQ_MTSLOTS;
  tab.tknull = (u32*)(tab.tkeys + SQ_MTSLOTS * (SQ_NKEYS > 0 ? SQ_NKEYS : 1));
  tab.tstate = tab.tknull + SQ_MTSLOTS;
  tab.tgroup = tab.tstate + SQ_MTSLOTS;
  tab.ngroups = tab.tgroup + SQ_MTSLOTS;
  u32* flags = tab.ngroups + 1;
  u64* thash = tab.thash;
  u64* tkeys = tab.tkeys;
  u32* tknull = tab.tknull;
  u32* tstate = tab.tstate;
  u32* tgroup = tab.tgroup;
  u32* ngroups = tab.ngroups;
  const int tid = threadIdx.x;
  if ((*((volatile u32*)status) & SQ_STATUS_OVERFLOW2) != 0u) return;

  for (int i = tid; i < SQ_ACC_WORDS * SQ_MSLOTS; i += 256) {
    const int w = i / SQ_MSLOTS;
    macc[i] = (w == SQ_NACC) ? SQ_EMPTY_ROW : sq_acc_identity(w);
  }
  for (int s = tid; s < SQ_MTSLOTS; s += 256) {
    tstate[s] = 0u;
    tab.ttag[s] = 0ULL;
  }
  if (tid == 0) {
    *ngroups = 0u;
    *flags = 0u;
  }
  __syncthreads();

  bool any_err = false, overflow = false;
  const i64 tile = (i64)256 * SQ_MUNROLL;
  for (i64 base = (i64)blockIdx.x * tile; base < n; base += (i64)gridDim.x * tile) {
    if (*((volatile u32*)flags) != 0u || (*((volatile u32*)status) & SQ_STATUS_OVERFLOW2) != 0u) break;  // some CTA overflowed: the batch is re-run anyway
    SqRow o[SQ_MUNROLL];
    bool live[SQ_MUNROLL];
#pragma unroll
    for (int u = 0; u < SQ_MUNROLL; u++) {
      const i64 r = base + (i64)u * 256 + tid;
      const bool inb = r < n;
      bool e0 = false, e1 = false;
      sq_row(in, inb ? r : n - 1, o[u], e0, e1);
      live[u] = inb && o[u].pass;
      any_err |= (inb && e0) || (live[u] && e1);
    }
    int g[SQ_MUNROLL];
    bool miss = false;
#pragma unroll
    for (int u = 0; u < SQ_MUNROLL; u++) {
      g[u] = live[u] ? sq_small_find<SQ_MTSLOTS>(tab, o[u]) : -1;
      miss |= live[u] && g[u] < 0;
    }
    if (__any_sync(SQ_FULL, miss)) {  // an unpublished key: out-of-line blocking protocol for this tile
      const u32 ret = sq_medium_tile_slow(in, n, row_base, base, macc, tab);
      any_err |= (ret & 1u) != 0u;
      if (ret & 2u) {
        overflow = true;
        *((volatile u32*)flags) = 1u;
      }
      continue;
    }
#pragma unroll
    for (int u = 0; u < SQ_MUNROLL; u++) {
      if (g[u] >= 0) {
        u64 local[SQ_NACC > 0 ? SQ_NACC : 1];
#pragma unroll
        for (int w = 0; w < SQ_NACC; w++) local[w] = sq_acc_identity(w);
        sq_acc_update(local, 1, o[u]);
#pragma unroll
        for (int w = 0; w < SQ_NACC; w++) sq_acc_merge_shared(&macc[(size_t)w * SQ_MSLOTS + g[u]], w, local[w]);
        atomicMin(&macc[(size_t)SQ_NACC * SQ_MSLOTS + g[u]], (u64)(row_base + base + (i64)u * 256 + tid));
      }
    }
    __syncwarp();
  }
  if (any_err) atomicOr(err, 1u);
  if (overflow) *((volatile u32*)flags) = 1u;
  __syncthreads();
  if (*((volatile u32*)flags) != 0u) {  // more than SQ_MSLOTS groups in this CTA: the host reruns the batch on sq_agg_global
    if (tid == 0) atomicOr(status, SQ_STATUS_OVERFLOW2);
    return;
  }
  if ((*((volatile u32*)status) & SQ_STATUS_OVERFLOW2) != 0u) return;  // another CTA overflowed: nothing of this batch counts
  // flush: the CTA's groups go to the scratch area; sq_agg_merge folds them into the HBM table only if NO
  // CTA overflowed (a flush straight into the table could not be undone before the re-run)
  for (int g = tid; g < SQ_MSLOTS; g += 256) part.state[(size_t)blockIdx.x * SQ_MSLOTS + g] = 0u;
  __syncthreads();
  for (int s = tid; s < SQ_MTSLOTS; s += 256) {
    if (tstate[s] != 2u) continue;
    const u32 g = tgroup[s];
    if (g >= SQ_MSLOTS) continue;
    const size_t e = (size_t)blockIdx.x * SQ_MSLOTS + g;
    part.state[e] = 2u;
    part.hash[e] = thash[s];
    part.knull[e] = tknull[s];
    part.min_row[e] = macc[(size_t)SQ_NACC * SQ_MSLOTS + g];
#pragma unroll
    for (int k = 0; k < SQ_NKEYS; k++) part.keys[e * SQ_NKEYS + k] = tkeys[s * (SQ_NKEYS > 0 ? SQ_NKEYS : 1) + k];
#pragma unroll
    for (int w = 0; w < SQ_NACC; w++) part.acc[e * SQ_NACC + w] = macc[(size_t)w * SQ_MSLOTS + g];
  }
}

// many groups: straight into the HBM table.  The table is kept DENSE: it starts small and a row whose group is new while the
// table holds `limit` groups is not inserted but appended to `overflow_rows`; the host grows the table (x4, re-hash) and
// launches the kernel again over that list (`redo_rows`).  Measured reason (profiles/r02_q1_sf10_variants.txt): sized for the
// worst case "every row a new group", 125 k groups scattered over 16 M slots x 13 arrays touch ~200 MB of L2 lines and every
// row pays DRAM latency a dozen times; in a 256 k-slot table the same groups are L2-resident.
// SQ_GUNROLL rows per thread per trip: their loads and their find-or-insert probes are independent chains, issued back to
// back; the first-row id only goes through an atomic when it lowers the slot's current value, the accumulator words are
// fire-and-forget reductions.
#ifndef SQ_GUNROLL
#define SQ_GUNROLL 2
#endif
extern "C" __global__ void __launch_bounds__(256) sq_agg_global(SqIn in, i64 n, i64 row_base, SqTable table, i64 batch_no, u32* __restrict__ status,
                                                                 u32* __restrict__ err, const u32* __restrict__ redo_rows, u32 n_redo,
                                                                 u32* __restrict__ overflow_rows, u32* __restrict__ overflow_count, u32 limit) {
  bool any_err = false;
  const i64 total = redo_rows ? (i64)n_redo : n;
  const i64 stride = (i64)gridDim.x * blockDim.x * SQ_GUNROLL;
  // warp-uniform trip count (the tail is predicated) so that __syncwarp can reconverge the warp
  // after the divergent find-or-insert
  for (i64 base = ((i64)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * SQ_GUNROLL; base < total; base += stride) {
    SqRow o[SQ_GUNROLL];
    bool live[SQ_GUNROLL];
    i64 row[SQ_GUNROLL];
#pragma unroll
    for (int u = 0; u < SQ_GUNROLL; u++) {
      const i64 i = base + u * 32 + (threadIdx.x & 31);
      const bool inb = i < total;
      row[u] = inb ? (redo_rows ? (i64)redo_rows[i] : i) : n - 1;
      bool e0 = false, e1 = false;
      sq_row(in, row[u], o[u], e0, e1);
      live[u] = inb && o[u].pass;
      any_err |= (inb && e0) || (live[u] && e1);
    }
    int slot[SQ_GUNROLL];
#pragma unroll
    for (int u = 0; u < SQ_GUNROLL; u++) {
      slot[u] = -1;
      if (live[u]) {
        slot[u] = sq_table_upsert(table, o[u].h, o[u].kb, o[u].knull, limit);
        if (slot[u] == -1) atomicOr(status, SQ_STATUS_FULL);
        if (slot[u] == -2) overflow_rows[atomicAdd(overflow_count, 1u)] = (u32)row[u];
      }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < SQ_GUNROLL; u++) {
      if (slot[u] >= 0) {
        const u64 gr = (u64)(row_base + row[u]);
        if (gr < table.min_row[slot[u]]) atomicMin(&table.min_row[slot[u]], gr);
        u64 local[SQ_NACC > 0 ? SQ_NACC : 1];
#pragma unroll
        for (int w = 0; w < SQ_NACC; w++) local[w] = sq_acc_identity(w);
        sq_acc_update(local, 1, o[u]);
#pragma unroll
        for (int w = 0; w < SQ_NACC; w++) sq_acc_merge_global(&table.acc[(size_t)w * table.capacity + slot[u]], w, local[w], batch_no);
      }
    }
    __syncwarp();
  }
  if (any_err) atomicOr(err, 1u);
}

// quirk K2 exactness: with hash-only identity a group may hold rows with different key tuples and
// the reference reports the keys of its FIRST row (hash_agg.rs:92-96).  After a batch, the slots
// inserted during it get the key bits of their minimum row.
extern "C" __global__ void __launch_bounds__(128) sq_agg_fixkeys(SqIn in, i64 n, i64 row_base, SqTable table, u32 n_new) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_new) return;
  const u32 slot = table.new_slots[i];
  const i64 r = (i64)(table.min_row[slot]) - row_base;
  if (r < 0 || r >= n) return;
  SqRow o;
  bool e0 = false, e1 = false;
  sq_row(in, r, o, e0, e1);
#pragma unroll
  for (int k = 0; k < SQ_NKEYS; k++) table.keys[(size_t)k * table.capacity + slot] = o.kb[k];
  table.knull[slot] = o.knull;
}
