// sqlrs_b200 JIT building block "join_table": the probe side's view of a sealed hash-join build table
// (kernels_join.cu: open-addressed slots over the distinct keys + CSR row lists + blocked Bloom filter) and its exact
// lookup, shared by the fused probe kernels (joinprobe.cuh, joinagg.cuh).  Needs SqProbe / SQ_JKEYS / SQ_JMATCH from
// the generated part.  Reference: the HashMap<u64, Vec<usize>> lookup of src/executor/join/hash_join.rs:225-235.

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
  const u32* bloom;       // blocked Bloom filter over the build hashes (3 bits in one 32-bit word per key)
  u32 bloom_mask;
  int unique;             // every build key occurs once: the match list of a slot is its representative row
  const u64* kv;          // key-in-slot layout (single key compared by value): kv[2s] = key bits, kv[2s+1] = representative row
  int kv_dtype;
  int rep_stride;         // slot_rep[slot * rep_stride] (2 in kv mode, where slot_rep = kv + 1)
  i64 n_inserted;
};
#define SQ_KV_EMPTY 0xffffffffffffffffULL
#ifndef SQ_KV_L2_SLOTS
#define SQ_KV_L2_SLOTS (4u << 20)  /* a kv table with more slots (> 64 MB) is not worth L2 space: its probes are single-use */
#endif

// (mirrors join_bloom_word / join_bloom_bits of kernels_aot.hpp: all 32-bit arithmetic on the two halves of the hash)
__device__ __forceinline__ u32 sq_bloom_word(u64 h, u32 mask) { return (u32)(h >> 32) & mask; }
__device__ __forceinline__ u32 sq_bloom_bits(u64 h) {
  const u32 lo = (u32)h;
  return (1u << (lo & 31u)) | (1u << ((lo >> 5) & 31u)) | (1u << ((lo >> 10) & 31u));
}

// exact lookup: the matched slot (or -1) and its representative build row
__device__ __forceinline__ int sq_join_find_rep(const SqJoin& t, const SqProbe& p, i64& rep_out) {
#if SQ_JMATCH
  if (p.knull != 0u) return -1;  // SQL semantics: a NULL key never joins
#endif
  const u32 mask = t.capacity - 1;
  u32 s = sq_mix32(p.h) & mask;
#if SQ_JMATCH && SQ_JKEYS == 1
  if (t.kv) {  // ONE 16-byte read per probe step: {key bits, representative row}
    if (t.kv_dtype != SQ_JKEY0_DTYPE || p.kb[0] == SQ_KV_EMPTY) return -1;
    for (u32 probes = 0; probes <= mask; probes++) {
      const ulonglong2 e = t.capacity > SQ_KV_L2_SLOTS ? sq_ld_u64x2_l2((const ulonglong2*)t.kv + s, sq_l2_evict_first()) : __ldg((const ulonglong2*)t.kv + s);
      if (e.x == p.kb[0]) {
        rep_out = (i64)e.y;
        return (int)s;
      }
      if (e.x == SQ_KV_EMPTY) return -1;
      s = (s + 1) & mask;
    }
    return -1;
  }
#endif
  for (u32 probes = 0; probes <= mask; probes++) {
    const i64 rep = __ldg(&t.slot_rep[s]);
    if (rep < 0) return -1;
    if (__ldg(&t.h[rep]) == p.h) {
#if SQ_JMATCH
      bool same = true;
#pragma unroll
      for (int k = 0; k < SQ_JKEYS; k++) same = same && (__ldg(&t.keys[(size_t)k * t.n_build + rep]) == p.kb[k]);
      if (same) {
        rep_out = rep;
        return (int)s;
      }
#else
      rep_out = rep;
      return (int)s;
#endif
    }
    s = (s + 1) & mask;
  }
  return -1;
}
__device__ __forceinline__ int sq_join_find(const SqJoin& t, const SqProbe& p) {
  i64 rep;
  return sq_join_find_rep(t, p, rep);
}

// phase A for SQ_JUNROLL x 32 rows starting at `base` (shared by the register-staged kernels): candidates are appended to
// the warp's queue (row number + the value phase B needs, see SQ_PQMODE)
#define SQ_PHASE_A(UNROLL, ROW_EXPR)                                                                          \
  {                                                                                                           \
    u64 qv[UNROLL];                                                                                           \
    u32 bits[UNROLL], bw[UNROLL];                                                                             \
    bool live[UNROLL];                                                                                        \
    _Pragma("unroll") for (int u = 0; u < UNROLL; u++) {                                                      \
      const i64 r = base + u * 32 + lane;                                                                     \
      const bool inb_row = r < n;                                                                             \
      SqProbe p;                                                                                              \
      bool e0 = false, e1 = false;                                                                            \
      sq_probe_row(in, inb_row ? r : n - 1, p, e0, e1);                                                       \
      live[u] = inb_row && p.pass;                                                                            \
      if (SQ_JMATCH) live[u] = live[u] && p.knull == 0u; /* SQL semantics: a NULL key never joins */          \
      any_err |= (inb_row && e0) || (inb_row && p.pass && e1);                                                \
      qv[u] = SQ_PQMODE ? sq_probe_qv(p) : 0ULL;                                                              \
      bits[u] = sq_bloom_bits(p.h);                                                                           \
      bw[u] = sq_bloom_word(p.h, jt.bloom_mask);                                                              \
    }                                                                                                         \
    _Pragma("unroll") for (int u = 0; u < UNROLL; u++) bw[u] = live[u] ? sq_ld_u32_l2(&jt.bloom[bw[u]], pol_keep) : 0u; \
    _Pragma("unroll") for (int u = 0; u < UNROLL; u++) {                                                      \
      const bool cand = live[u] && (bw[u] & bits[u]) == bits[u];                                              \
      const u32 m = __ballot_sync(0xffffffffu, cand);                                                         \
      if (cand) {                                                                                             \
        const u32 pos = queued + __popc(m & lanes_below);                                                     \
        queue[pos] = (u32)(ROW_EXPR);                                                                         \
        if (SQ_PQMODE) queue_v[pos] = qv[u];                                                                  \
      }                                                                                                       \
      queued += __popc(m);                                                                                    \
    }                                                                                                         \
    __syncwarp();                                                                                             \
  }
#if !SQ_PQMODE
__device__ __forceinline__ u64 sq_probe_qv(const SqProbe&) { return 0ULL; }
#endif

