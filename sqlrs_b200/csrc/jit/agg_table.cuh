// sqlrs_b200 JIT building block "agg_table": the operator's group table in HBM and its find-or-insert, shared by
// the aggregate skeletons (agg.cuh, joinagg.cuh).  Needs SQ_NKEYS and SQ_MATCH_KEYS from the generated part.
struct SqTable {      // the operator's persistent group table in HBM (SoA, capacity = power of two)
  u32* state;         // 0 empty, 1 being written, 2 ready
  u64* hash;          // group identity (row hash)
  u64* min_row;       // first global row id of the group
  u64* keys;          // [K][capacity] raw key bits (NULL cells 0)
  u32* knull;         // null mask of the key tuple
  u64* acc;           // [W][capacity]
  u32* new_slots;     // slots inserted since the last key fix-up
  u32* counters;      // [0] groups, [1] new_slots length, [2] status bits
  u32 capacity;
};
#define SQ_STATUS_OVERFLOW 1u   /* sq_agg_small: more groups than SQ_SLOTS in some CTA */
#define SQ_STATUS_FULL 2u       /* table ran out of slots (host sized it wrongly) */
#define SQ_STATUS_OVERFLOW2 4u  /* sq_agg_medium: more groups than SQ_MSLOTS in some CTA */

struct SqPartial {    // CTA partials of sq_agg_small: [cta][slot]
  u32* state;
  u64* hash;
  u64* min_row;
  u64* keys;          // [(cta*S+slot)*K + k]
  u32* knull;
  u64* acc;           // [(cta*S+slot)*W + w]
};

__device__ __forceinline__ bool sq_keys_equal(const u64* a, u32 an, const u64* b, u32 bn) {
#if SQ_MATCH_KEYS
  if (an != bn) return false;
#pragma unroll
  for (int k = 0; k < SQ_NKEYS; k++)
    if (a[k] != b[k]) return false;
#endif
  return true;
}

// Publication protocol of a slot: the publisher writes hash / keys / null mask, then stores state = 2 with release semantics.
// A reader loads the state from L2 (volatile) and — only after it saw 2: the loads are control-dependent — the identity
// words with ld.global.cg, i.e. from L2 as well, never from a possibly stale L1 line.  No fence and no acquire on the read
// side: __threadfence() per lookup is a MEMBAR.GPU per row, ld.acquire.gpu compiles to an L1 invalidate (CCTL.IVALL, 20 % of
// the stall samples of sq_agg_global in profiles/r02f_*).
__device__ __forceinline__ u32 sq_ld_acquire_u32(const u32* p) { return *((volatile const u32*)p); }
__device__ __forceinline__ void sq_st_release_u32(u32* p, u32 v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// find-or-insert in the HBM table; returns the slot, -1 (table full), or -2: the key is absent and the table already holds
// `limit` groups (the caller keeps the table dense: it records the row, grows the table and redoes the row)
__device__ __forceinline__ int sq_table_upsert(const SqTable& t, u64 h, const u64* kb, u32 knull, u32 limit = 0xffffffffu) {
  const u32 mask = t.capacity - 1;
  u32 s = sq_mix32(h) & mask;
  for (u32 probes = 0; probes <= mask;) {
    u32 st = sq_ld_acquire_u32(&t.state[s]);
    if (st == 0) {
      if (limit != 0xffffffffu && *((volatile u32*)&t.counters[0]) >= limit) return -2;
      if (atomicCAS(&t.state[s], 0u, 1u) == 0u) {
        t.hash[s] = h;
#pragma unroll
        for (int k = 0; k < SQ_NKEYS; k++) t.keys[(size_t)k * t.capacity + s] = kb[k];
        t.knull[s] = knull;
        sq_st_release_u32(&t.state[s], 2u);
        atomicAdd(&t.counters[0], 1u);
        t.new_slots[atomicAdd(&t.counters[1], 1u)] = s;
        return (int)s;
      }
      continue;  // lost the race: look at the slot again
    }
    if (st == 1) continue;  // another thread is publishing this slot
    if (__ldcg(&t.hash[s]) == h) {   // published slots never change their identity
#if SQ_MATCH_KEYS
      u64 other[SQ_NKEYS > 0 ? SQ_NKEYS : 1];
#pragma unroll
      for (int k = 0; k < SQ_NKEYS; k++) other[k] = __ldcg(&t.keys[(size_t)k * t.capacity + s]);
      if (sq_keys_equal(other, __ldcg(&t.knull[s]), kb, knull)) return (int)s;
#else
      return (int)s;
#endif
    }
    s = (s + 1) & mask;
    probes++;
  }
  return -1;
}
