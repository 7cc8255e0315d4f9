// sqlrs_b200 JIT skeleton "joinbuild": fused  scan -> Filter -> join-key hash -> INSERT of one build batch into a key-in-slot
// table (JoinTableView::kv) + its Bloom filter.  Replaces, for a single compared key, the key-evaluation pass (hash / raw key /
// null mask / keep bitmap written to HBM), the bit count of the keep bitmap and the insert kernel reading them back
// (src/executor/join/hash_join.rs:161-181 builds a HashMap<u64, Vec<usize>> row by row): the build columns are read ONCE.
// Generated in front of this file: SqIn, SqProbe + sq_probe_row (the Filter fused below the build side + the LEFT key
// expressions, compiled with the same generator as a probe side's), SQ_JKEYS == 1, SQ_JMATCH == 1; then join_table.cuh.
// out.kv == nullptr: count-only pass (the host sizes the table from the number of rows the Filter keeps).
#define SQ_BBLOCK 256

struct SqBuildOut {
  u64* kv;      // [2 * capacity], initialised to SQ_KV_EMPTY
  u32* bloom;
  u32 capacity;
  u32 bloom_mask;
  u32* flags;   // [0] some key repeats, [3] expression error, [4] a key equals the empty marker, [5] table full
  u64* kept;    // rows that pass the Filter with a non-NULL key
};

// SQ_BWAY queued rows per lane are inserted together: all first CAS are issued before the first result is used
#define SQ_BWAY 4
#define SQ_BQUEUE (SQ_BUNROLL8 * 32 + 32 * SQ_BWAY)
#define SQ_BUNROLL8 8

__device__ __forceinline__ void sq_build_insert_tail(u64 c, u32 s, u64 key, u64 h, u64 row, const SqBuildOut& out, bool& dup, bool& full) {
  const u32 mask = out.capacity - 1;
  for (u32 probes = 0;; probes++) {
    if (c == SQ_KV_EMPTY) break;  // claimed
    if (c == key) {
      dup = true;
      break;
    }
    if (probes >= mask) {
      full = true;
      return;
    }
    s = (s + 1) & mask;
    c = atomicCAS(&out.kv[2 * (size_t)s], SQ_KV_EMPTY, key);
  }
  atomicMin(&out.kv[2 * (size_t)s + 1], row);  // the representative row = the key's first build row
  sq_red_or_u32_l2(&out.bloom[sq_bloom_word(h, out.bloom_mask)], sq_bloom_bits(h), sq_l2_evict_last());
}

// Rows the Filter keeps are usually a fraction of the scan (Q3': 20 %): they are compacted into a per-warp queue (ballot / popc)
// and inserted by FULL warps, SQ_BWAY per lane at a time — the random atomics never run with mostly idle lanes.
extern "C" __global__ void __launch_bounds__(SQ_BBLOCK) sq_joinbuild_kernel(SqIn in, i64 n, SqBuildOut out) {
  __shared__ u32 queue_s[SQ_BBLOCK / 32][SQ_BQUEUE];
  __shared__ u64 queue_vs[SQ_BBLOCK / 32][SQ_BQUEUE];
  const u32 mask = out.capacity - 1;
  const int lane = threadIdx.x & 31;
  u32* queue = queue_s[threadIdx.x >> 5];
  u64* queue_v = queue_vs[threadIdx.x >> 5];
  const u32 lanes_below = (1u << lane) - 1u;
  bool any_err = false, dup = false, sentinel = false, full = false;
  u32 kept = 0, queued = 0;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 base = ((i64)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * SQ_BUNROLL8; base < n; base += stride * SQ_BUNROLL8) {
    u64 key[SQ_BUNROLL8];
    bool live[SQ_BUNROLL8];
#pragma unroll
    for (int u = 0; u < SQ_BUNROLL8; u++) {
      const i64 r = base + u * 32 + lane;
      const bool inb = r < n;
      SqProbe p;
      bool e0 = false, e1 = false;
      sq_probe_row(in, inb ? r : n - 1, p, e0, e1);
      live[u] = inb && p.pass && p.knull == 0u;  // SQL semantics: a NULL key never joins
      any_err |= (inb && e0) || (inb && p.pass && e1);
      key[u] = p.kb[0];
      if (live[u] && key[u] == SQ_KV_EMPTY) {
        sentinel = true;
        live[u] = false;
      }
      kept += live[u] ? 1u : 0u;
    }
    if (!out.kv) continue;
#pragma unroll
    for (int u = 0; u < SQ_BUNROLL8; u++) {
      const u32 m = __ballot_sync(0xffffffffu, live[u]);
      if (live[u]) {
        const u32 pos = queued + __popc(m & lanes_below);
        queue[pos] = (u32)(base + u * 32 + lane);  // n < 2^32 (checked by the host)
        queue_v[pos] = key[u];
      }
      queued += __popc(m);
    }
    __syncwarp();
    while (queued >= 32 * SQ_BWAY) {
      queued -= 32 * SQ_BWAY;
      u64 k[SQ_BWAY], h[SQ_BWAY], c[SQ_BWAY];
      u32 row[SQ_BWAY];
#pragma unroll
      for (int w = 0; w < SQ_BWAY; w++) {
        row[w] = queue[queued + w * 32 + lane];
        k[w] = queue_v[queued + w * 32 + lane];
        const u64 kk[1] = {k[w]};
        h[w] = sq_probe_rehash(kk);
      }
#pragma unroll
      for (int w = 0; w < SQ_BWAY; w++) c[w] = atomicCAS(&out.kv[2 * (size_t)(sq_mix32(h[w]) & mask)], SQ_KV_EMPTY, k[w]);
#pragma unroll
      for (int w = 0; w < SQ_BWAY; w++) sq_build_insert_tail(c[w], sq_mix32(h[w]) & mask, k[w], h[w], (u64)row[w], out, dup, full);
      __syncwarp();
    }
  }
  // the rest of the queue, one row per lane at a time
  for (u32 i = lane; i < queued; i += 32) {
    const u64 k = queue_v[i];
    const u64 kk[1] = {k};
    const u64 h = sq_probe_rehash(kk);
    const u32 s = sq_mix32(h) & mask;
    sq_build_insert_tail(atomicCAS(&out.kv[2 * (size_t)s], SQ_KV_EMPTY, k), s, k, h, (u64)queue[i], out, dup, full);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, d);
  if (lane == 0 && kept) atomicAdd(out.kept, (u64)kept);
  if (any_err) atomicOr(&out.flags[3], 1u);
  if (dup) atomicOr(&out.flags[0], 1u);
  if (sentinel) atomicOr(&out.flags[4], 1u);
  if (full) atomicOr(&out.flags[5], 1u);
}
