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

// find-or-insert in the HBM table; returns the slot or -1 (table full)
__device__ __forceinline__ int sq_table_upsert(const SqTable& t, u64 h, const u64* kb, u32 knull) {
  const u32 mask = t.capacity - 1;
  u32 s = sq_mix32(h) & mask;
  for (u32 probes = 0; probes <= mask;) {
    u32 st = *((volatile u32*)&t.state[s]);
    if (st == 0) {
      if (atomicCAS(&t.state[s], 0u, 1u) == 0u) {
        t.hash[s] = h;
#pragma unroll
        for (int k = 0; k < SQ_NKEYS; k++) t.keys[(size_t)k * t.capacity + s] = kb[k];
        t.knull[s] = knull;
        __threadfence();
        atomicExch(&t.state[s], 2u);
        atomicAdd(&t.counters[0], 1u);
        t.new_slots[atomicAdd(&t.counters[1], 1u)] = s;
        return (int)s;
      }
      continue;  // lost the race: look at the slot again
    }
    if (st == 1) continue;  // another thread is publishing this slot
    __threadfence();
    if (*((volatile u64*)&t.hash[s]) == h) {
#if SQ_MATCH_KEYS
      u64 other[SQ_NKEYS > 0 ? SQ_NKEYS : 1];
#pragma unroll
      for (int k = 0; k < SQ_NKEYS; k++) other[k] = *((volatile u64*)&t.keys[(size_t)k * t.capacity + s]);
      if (sq_keys_equal(other, *((volatile u32*)&t.knull[s]), kb, knull)) return (int)s;
#else
      return (int)s;
#endif
    }
    s = (s + 1) & mask;
    probes++;
  }
  return -1;
}

