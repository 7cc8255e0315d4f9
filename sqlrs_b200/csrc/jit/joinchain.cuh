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
#define SQ_CUNROLL 8
#define SQ_CQUEUE (SQ_CUNROLL * 32 + 32)

struct SqChainOut {
  u64* kv;        // [2 * capacity], initialised to SQ_KV_EMPTY
  u64* bloom;
  u32 capacity;   // power of two
  u32 bloom_mask;
  u32* flags;     // [0] some key repeats, [1] table full, [2] a key equals the empty marker, [3] expression error
  u64* inserted;  // number of rows inserted
};

__device__ __forceinline__ u32 sq_chain_candidate(const SqIn& in, const SqInB& inb, i64 r, const SqJoin& jt, const SqChainOut& out, bool& any_err) {
  SqProbe p;
  bool e0 = false, e1 = false;
  sq_probe_row(in, r, p, e0, e1);
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
  const u32 mask = out.capacity - 1;
  u32 s = sq_mix32(k.h) & mask;
  u64 cur = *((volatile u64*)&out.kv[2 * (size_t)s]);
  u32 probes = 0;
  for (;;) {
    if (cur == SQ_KV_EMPTY) cur = atomicCAS(&out.kv[2 * (size_t)s], SQ_KV_EMPTY, k.kb);
    if (cur == SQ_KV_EMPTY) break;  // claimed
    if (cur == k.kb) {
      atomicOr(&out.flags[0], 1u);
      break;
    }
    if (++probes > mask) {
      atomicOr(&out.flags[1], 1u);
      return 0u;
    }
    s = (s + 1) & mask;
    cur = *((volatile u64*)&out.kv[2 * (size_t)s]);
  }
  atomicMin(&out.kv[2 * (size_t)s + 1], (u64)r);
  atomicOr(&out.bloom[sq_bloom_word(k.h, out.bloom_mask)], sq_bloom_bits(k.h));
  return 1u;
}

// chunk_step > 1: only every chunk_step-th 2048-row chunk is processed (sampling, with out.kv == nullptr)
extern "C" __global__ void __launch_bounds__(SQ_CBLOCK) sq_joinchain_kernel(SqIn in, SqInB inb, i64 n, SqJoin jt, SqChainOut out, i64 chunk_step) {
  __shared__ u32 queue_s[SQ_CBLOCK / 32][SQ_CQUEUE];
  bool any_err = false;
  const int lane = threadIdx.x & 31;
  u32* queue = queue_s[threadIdx.x >> 5];
  const u32 lanes_below = (1u << lane) - 1u;
  u32 queued = 0;    // warp-uniform
  u32 inserted = 0;  // per lane
  for (i64 trip = blockIdx.x;; trip += gridDim.x) {
    const i64 base = trip * chunk_step * (SQ_CBLOCK * SQ_CUNROLL) + (i64)(threadIdx.x & ~31) * SQ_CUNROLL;
    if (base >= n) break;
    // ---- phase A: streaming Filter + key hash + Bloom test of join 1 (see joinagg.cuh)
    u64 hh[SQ_CUNROLL];
    bool live[SQ_CUNROLL];
#pragma unroll
    for (int u = 0; u < SQ_CUNROLL; u++) {
      const i64 r = base + u * 32 + lane;
      const bool inb_row = r < n;
      SqProbe p;
      bool e0 = false, e1 = false;
      sq_probe_row(in, inb_row ? r : n - 1, p, e0, e1);
      live[u] = inb_row && p.pass;
#if SQ_JMATCH
      live[u] = live[u] && p.knull == 0u;
#endif
      any_err |= (inb_row && e0) || (inb_row && p.pass && e1);
      hh[u] = p.h;
    }
    u64 bw[SQ_CUNROLL];
#pragma unroll
    for (int u = 0; u < SQ_CUNROLL; u++) bw[u] = live[u] ? __ldg(&jt.bloom[sq_bloom_word(hh[u], jt.bloom_mask)]) : 0ULL;
#pragma unroll
    for (int u = 0; u < SQ_CUNROLL; u++) {
      const u64 bits = sq_bloom_bits(hh[u]);
      const bool cand = live[u] && (bw[u] & bits) == bits;
      const u32 m = __ballot_sync(0xffffffffu, cand);
      if (cand) queue[queued + __popc(m & lanes_below)] = (u32)(base + u * 32 + lane);  // n < 2^32 (checked by the host)
      queued += __popc(m);
    }
    __syncwarp();
    // ---- phase B: full warps probe join 1's table and insert into join 2's
    while (queued >= 32) {
      queued -= 32;
      inserted += sq_chain_candidate(in, inb, (i64)queue[queued + lane], jt, out, any_err);
      __syncwarp();
    }
  }
  if ((u32)lane < queued) inserted += sq_chain_candidate(in, inb, (i64)queue[lane], jt, out, any_err);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) inserted += __shfl_xor_sync(0xffffffffu, inserted, d);
  if (lane == 0 && inserted) atomicAdd(out.inserted, (u64)inserted);
  if (any_err) atomicOr(&out.flags[3], 1u);
}
