// sqlrs_b200 JIT skeleton "joinchain": fused  scan -> Filter -> hash-join PROBE (join 1) -> hash-join BUILD (join 2).
// For a left-deep chain  (A join B) join C  the reference materialises the whole output of the first join
// (build_batch gathers of every column, src/executor/join/hash_join.rs:25-45,284-291), concatenates it (:187) and
// re-hashes it into the second join's HashMap (:161-181).  Here the probe of join 1 inserts each matching row's
// join-2 key straight into join 2's key-in-slot table (JoinTableView::kv): no (build row, probe row) pairs, no
// gathered columns, no second pass.  The virtual build row of join 2 is the PROBE row id r of join 1 — join 1's build keys
// are unique (checked by the host), so probe row r yields at most one joined row and r ascends with the reference's
// output order of join 1 (probe-row order, :225-248); join 2's payload columns are read in place from B at row r.
//
// Generated in front of this file: SqIn (B), SqInB (A), SqProbe + sq_probe_row (Filter below join 1 on the probe side + join-1
// keys), SQ_JKEYS / SQ_JMATCH / SQ_JKEY0_DTYPE, and
//   struct SqChainKey {u64 h; u64 kb; u32 knull;};  sq_chain_key(in, inb, r, b, k, e)  — join 2's build key over the joined row
// then join_table.cuh.  HBM-bound on the scan of B's Filter/key columns; the inserts are random 16-byte atomics that overlap it.
#define SQ_CBLOCK 256
#ifndef SQ_CUNROLL
#define SQ_CUNROLL 8
#endif
#ifndef SQ_CWAY
#define SQ_CWAY 4  // candidates per lane in one phase-B pass (their dependent chains are interleaved stage by stage)
#endif
#ifndef SQ_PREFETCH
#define SQ_PREFETCH 0  // L2 prefetch distance in trips (0 = off)
#endif
#ifndef SQ_CMINB
#define SQ_CMINB 4  // __launch_bounds__ minimum CTAs per SM: 64 registers, 32 resident warps (72 registers / 24 warps unconstrained:
#endif              // 6 % slower on the whole Q3' query at SF100, profiles/r02h_q3_knobs.txt)
#define SQ_CQUEUE (SQ_CUNROLL * 32 + 32 * SQ_CWAY)

struct SqChainOut {
  u64* kv;        // [2 * capacity], initialised to SQ_KV_EMPTY
  u32* bloom;
  u32 capacity;   // power of two
  u32 bloom_mask;
  u32* flags;     // [0] some key repeats, [1] table full, [2] a key equals the empty marker, [3] expression error
  u64* inserted;  // number of rows inserted
};

// the rest of an insert whose first CAS at slot `s` returned `cur` (collision chain), then the representative row + Bloom bits
__device__ __forceinline__ u32 sq_chain_insert_tail(u64 cur, u32 s, const SqChainKey& k, i64 r, const SqChainOut& out) {
  const u32 mask = out.capacity - 1;
  for (u32 probes = 0;; probes++) {
    if (cur == SQ_KV_EMPTY) {
      sq_st_u64_l2(&out.kv[2 * (size_t)s + 1], (u64)r, sq_l2_evict_first());
      break;
    }
    if (cur == k.kb) {
      atomicOr(&out.flags[0], 1u);
      return 0u;
    }
    if (probes >= mask) {
      atomicOr(&out.flags[1], 1u);
      return 0u;
    }
    s = (s + 1) & mask;
    cur = sq_cas_u64_l2(&out.kv[2 * (size_t)s], SQ_KV_EMPTY, k.kb, sq_l2_evict_first());
  }
  sq_red_or_u32_l2(&out.bloom[sq_bloom_word(k.h, out.bloom_mask)], sq_bloom_bits(k.h), sq_l2_evict_last());
  return 1u;
}

__device__ __forceinline__ u32 sq_chain_candidate(const SqIn& in, const SqInB& inb, i64 r, u64 qv, const SqJoin& jt, const SqChainOut& out, bool& any_err) {
  SqProbe p;
#if SQ_PQMODE
  sq_probe_unq(qv, p);
#else
  bool e0 = false, e1 = false;
  sq_probe_row(in, r, p, e0, e1);
#endif
  i64 b = -1;
  if (sq_join_find_rep(jt, p, b) < 0) return 0u;
  SqChainKey k;
  bool e2 = false;
  sq_chain_key(in, inb, r, b, k, e2);
  any_err |= e2;
  if (k.knull != 0u) return 0u;  // SQL semantics: a NULL key never joins
  if (!out.kv) return 1u;        // count-only pass (the host samples chunks to size the table)
  if (k.kb == SQ_KV_EMPTY) {
    atomicOr(&out.flags[2], 1u);
    return 0u;
  }
  // claim by CAS straight away (at load <= 0.5 the home slot is empty more often than not: ONE round trip to the slot's
  // line); the claimer alone writes the representative row with a plain store — a second row with the same key only
  // raises the flag, and a flagged table is discarded by the host
  const u32 s = sq_mix32(k.h) & (out.capacity - 1);
  return sq_chain_insert_tail(sq_cas_u64_l2(&out.kv[2 * (size_t)s], SQ_KV_EMPTY, k.kb, sq_l2_evict_first()), s, k, r, out);
}

#if SQ_PQMODE == 1 && !SQ_CHAIN_KEY_USES_BUILD
// Phase B for SQ_CWAY candidates per lane.  One candidate's chain is: probe join 1's table (1+ dependent reads) -> load
// join 2's key -> CAS into join 2's table (1+ dependent atomics).  Done one candidate per lane, a warp pays the LONGEST
// chain of its 32 lanes in full memory round trips (measured: 70 % of the kernel's stall samples sat on these waits).
// Here every stage is issued for all SQ_CWAY candidates before the first result is used: the first probes and the key
// loads (which do not depend on the probe's outcome: the key reads probe-side columns only) travel together, then all
// first CAS; only collisions continue one at a time.  Needs the key-in-slot layout for join 1's table.
__device__ __forceinline__ u32 sq_chain_batch(const SqIn& in, const SqInB& inb, const u32* rq, const u64* vq, int lane, const SqJoin& jt,
                                              const SqChainOut& out, bool& any_err) {
  const u32 mask1 = jt.capacity - 1, mask2 = out.capacity - 1;
  i64 r[SQ_CWAY];
  u64 kb1[SQ_CWAY];
  u32 s1[SQ_CWAY];
  ulonglong2 e[SQ_CWAY];
  SqChainKey k[SQ_CWAY];
  bool ok[SQ_CWAY];
#pragma unroll
  for (int u = 0; u < SQ_CWAY; u++) {
    r[u] = (i64)rq[u * 32 + lane];
    kb1[u] = vq[u * 32 + lane];
    const u64 kk[1] = {kb1[u]};
    s1[u] = sq_mix32(sq_probe_rehash(kk)) & mask1;
  }
#pragma unroll
  for (int u = 0; u < SQ_CWAY; u++) e[u] = __ldg((const ulonglong2*)jt.kv + s1[u]);
#pragma unroll
  for (int u = 0; u < SQ_CWAY; u++) {
    bool e2 = false;
    sq_chain_key(in, inb, r[u], 0, k[u], e2);
    ok[u] = !e2;  // an error only counts on a row that joins (below)
    if (e2) k[u].knull = 1u;
  }
  // resolve the probes of join 1 (collision chains one at a time)
#pragma unroll
  for (int u = 0; u < SQ_CWAY; u++) {
    bool found = jt.kv_dtype == SQ_JKEY0_DTYPE && kb1[u] != SQ_KV_EMPTY;
    if (found) {
      u32 s = s1[u];
      ulonglong2 cur = e[u];
      found = false;
      for (u32 probes = 0; probes <= mask1; probes++) {
        if (cur.x == kb1[u]) {
          found = true;
          break;
        }
        if (cur.x == SQ_KV_EMPTY) break;
        s = (s + 1) & mask1;
        cur = __ldg((const ulonglong2*)jt.kv + s);
      }
    }
    if (found && !ok[u]) any_err = true;
    ok[u] = found && k[u].knull == 0u;  // SQL semantics: a NULL key never joins
  }
  u32 done = 0;
  if (!out.kv) {  // count-only pass
#pragma unroll
    for (int u = 0; u < SQ_CWAY; u++) done += ok[u] ? 1u : 0u;
    return done;
  }
  u64 cur[SQ_CWAY];
  u32 s2[SQ_CWAY];
#pragma unroll
  for (int u = 0; u < SQ_CWAY; u++) {
    if (ok[u] && k[u].kb == SQ_KV_EMPTY) {
      atomicOr(&out.flags[2], 1u);
      ok[u] = false;
    }
    s2[u] = sq_mix32(k[u].h) & mask2;
  }
#pragma unroll
  for (int u = 0; u < SQ_CWAY; u++)
    if (ok[u]) cur[u] = sq_cas_u64_l2(&out.kv[2 * (size_t)s2[u]], SQ_KV_EMPTY, k[u].kb, sq_l2_evict_first());
#pragma unroll
  for (int u = 0; u < SQ_CWAY; u++)
    if (ok[u]) done += sq_chain_insert_tail(cur[u], s2[u], k[u], r[u], out);
  return done;
}
#define SQ_CHAIN_BATCHED 1
#else
#define SQ_CHAIN_BATCHED 0
#endif

// chunk_step > 1: only every chunk_step-th chunk (SQ_CBLOCK x SQ_CUNROLL rows) is processed (sampling, with out.kv == nullptr)
extern "C" __global__ void __launch_bounds__(SQ_CBLOCK, SQ_CMINB) sq_joinchain_kernel(SqIn in, SqInB inb, i64 n, SqJoin jt, SqChainOut out, i64 chunk_step) {
  __shared__ u32 queue_s[SQ_CBLOCK / 32][SQ_CQUEUE];
  __shared__ u64 queue_vs[SQ_CBLOCK / 32][SQ_PQMODE ? SQ_CQUEUE : 1];
  bool any_err = false;
  const int lane = threadIdx.x & 31;
  u32* queue = queue_s[threadIdx.x >> 5];
  u64* queue_v = queue_vs[threadIdx.x >> 5];
  const u32 lanes_below = (1u << lane) - 1u;
  u32 queued = 0;    // warp-uniform
  u32 inserted = 0;  // per lane
  const u64 pol_keep = sq_l2_evict_last();
  for (i64 trip = blockIdx.x;; trip += gridDim.x) {
    const i64 base = trip * chunk_step * (SQ_CBLOCK * SQ_CUNROLL) + (i64)(threadIdx.x & ~31) * SQ_CUNROLL;
    if (base >= n) break;
#if SQ_PREFETCH
    sq_probe_prefetch(in, base + SQ_PREFETCH * (i64)gridDim.x * chunk_step * (SQ_CBLOCK * SQ_CUNROLL), n, SQ_CUNROLL * 32, lane);
#endif
    // ---- phase A: streaming Filter + key hash + Bloom test of join 1 (see joinagg.cuh)
    SQ_PHASE_A(SQ_CUNROLL, base + u * 32 + lane)  // n < 2^32 (checked by the host)
    // ---- phase B: full warps probe join 1's table and insert into join 2's
#if SQ_CHAIN_BATCHED
    if (jt.kv) {
      while (queued >= 32 * SQ_CWAY) {
        queued -= 32 * SQ_CWAY;
        inserted += sq_chain_batch(in, inb, queue + queued, queue_v + queued, lane, jt, out, any_err);
        __syncwarp();
      }
      continue;
    }
#endif
    while (queued >= 32) {
      queued -= 32;
      inserted += sq_chain_candidate(in, inb, (i64)queue[queued + lane], SQ_PQMODE ? queue_v[queued + lane] : 0ULL, jt, out, any_err);
      __syncwarp();
    }
  }
  while (queued >= 32) {  // what the batched passes left over
    queued -= 32;
    inserted += sq_chain_candidate(in, inb, (i64)queue[queued + lane], SQ_PQMODE ? queue_v[queued + lane] : 0ULL, jt, out, any_err);
    __syncwarp();
  }
  if ((u32)lane < queued) inserted += sq_chain_candidate(in, inb, (i64)queue[lane], SQ_PQMODE ? queue_v[lane] : 0ULL, jt, out, any_err);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) inserted += __shfl_xor_sync(0xffffffffu, inserted, d);
  if (lane == 0 && inserted) atomicAdd(out.inserted, (u64)inserted);
  if (any_err) atomicOr(&out.flags[3], 1u);
}
