// sqlrs_b200 JIT skeleton "joinbuild": fused  scan -> Filter -> join-key hash -> INSERT of one build batch into a key-in-slot
// table (JoinTableView::kv) + its Bloom filter.  Replaces, for a single compared key, the key-evaluation pass (hash / raw key /
// null mask / keep bitmap written to HBM), the bit count of the keep bitmap and the insert kernel reading them back
// (src/executor/join/hash_join.rs:161-181 builds a HashMap<u64, Vec<usize>> row by row): the build columns are read ONCE.
// Generated in front of this file: SqIn, SqProbe + sq_probe_row (the Filter fused below the build side + the LEFT key
// expressions, compiled with the same generator as a probe side's), SQ_JKEYS == 1, SQ_JMATCH == 1; then join_table.cuh.
// out.kv == nullptr: count-only pass (the host sizes the table from the number of rows the Filter keeps).
#define SQ_BBLOCK 256
#define SQ_BUNROLL 4

struct SqBuildOut {
  u64* kv;      // [2 * capacity], initialised to SQ_KV_EMPTY
  u32* bloom;
  u32 capacity;
  u32 bloom_mask;
  u32* flags;   // [0] some key repeats, [3] expression error, [4] a key equals the empty marker, [5] table full
  u64* kept;    // rows that pass the Filter with a non-NULL key
};

extern "C" __global__ void __launch_bounds__(SQ_BBLOCK) sq_joinbuild_kernel(SqIn in, i64 n, SqBuildOut out) {
  const u32 mask = out.capacity - 1;
  bool any_err = false, dup = false, sentinel = false, full = false;
  u32 kept = 0;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 base = (i64)blockIdx.x * blockDim.x + threadIdx.x; base < n; base += stride * SQ_BUNROLL) {
    u64 key[SQ_BUNROLL], h[SQ_BUNROLL], cur[SQ_BUNROLL];
    bool live[SQ_BUNROLL];
#pragma unroll
    for (int u = 0; u < SQ_BUNROLL; u++) {
      const i64 r = base + u * stride;
      const bool inb = r < n;
      SqProbe p;
      bool e0 = false, e1 = false;
      sq_probe_row(in, inb ? r : n - 1, p, e0, e1);
      live[u] = inb && p.pass && p.knull == 0u;  // SQL semantics: a NULL key never joins
      any_err |= (inb && e0) || (inb && p.pass && e1);
      key[u] = p.kb[0];
      h[u] = p.h;
      if (live[u] && key[u] == SQ_KV_EMPTY) {
        sentinel = true;
        live[u] = false;
      }
      kept += live[u] ? 1u : 0u;
    }
    if (!out.kv) continue;
#pragma unroll
    for (int u = 0; u < SQ_BUNROLL; u++)
      if (live[u]) cur[u] = atomicCAS(&out.kv[2 * (size_t)(sq_mix32(h[u]) & mask)], SQ_KV_EMPTY, key[u]);
#pragma unroll
    for (int u = 0; u < SQ_BUNROLL; u++) {
      if (!live[u]) continue;
      u32 s = sq_mix32(h[u]) & mask;
      u64 c = cur[u];
      bool placed = true;
      for (u32 probes = 0;; probes++) {
        if (c == SQ_KV_EMPTY) break;  // claimed
        if (c == key[u]) {
          dup = true;
          break;
        }
        if (probes >= mask) {
          full = true;
          placed = false;
          break;
        }
        s = (s + 1) & mask;
        c = atomicCAS(&out.kv[2 * (size_t)s], SQ_KV_EMPTY, key[u]);
      }
      if (!placed) continue;
      atomicMin(&out.kv[2 * (size_t)s + 1], (u64)(base + u * stride));  // the representative row = the key's first build row
      sq_red_or_u32_l2(&out.bloom[sq_bloom_word(h[u], out.bloom_mask)], sq_bloom_bits(h[u]), sq_l2_evict_last());
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, d);
  if ((threadIdx.x & 31) == 0 && kept) atomicAdd(out.kept, (u64)kept);
  if (any_err) atomicOr(&out.flags[3], 1u);
  if (dup) atomicOr(&out.flags[0], 1u);
  if (sentinel) atomicOr(&out.flags[4], 1u);
  if (full) atomicOr(&out.flags[5], 1u);
}
