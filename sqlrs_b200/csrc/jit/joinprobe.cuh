// sqlrs_b200 JIT skeleton "joinprobe": fused  scan -> Filter -> join-key hash -> Bloom test -> exact PROBE of one
// probe batch.  Replaces, for every join type, the key-evaluation pass (hash / key / null-mask / keep columns
// written to HBM), the standalone probe-count kernel reading them back and the per-row offset scan of the
// reference-shaped pipeline (src/executor/join/hash_join.rs:208-248): the probe columns are read ONCE and the only
// per-row output is the matched slot (4 B).  Per 2048-row chunk the kernel also leaves the number of output rows,
// so that a scan over n/2048 chunk counts (not over n rows) positions the (build row, probe row) pairs, which
// k_join_probe_emit then writes in the reference's order (probe-row order, build insertion order per probe row).
//
// Generated in front of this file: SqIn, SqProbe, sq_probe_row (fused probe-side Filter + join keys), SQ_JKEYS, SQ_JMATCH;
// then join_table.cuh.
// slot_of[r]: >= 0 matched slot; -1 kept by the Filter but unmatched (Right/Full joins emit (NULL, r)); -2 dropped.
// HBM-bound: 8 B x referenced probe columns in + 4 B out per probe row; Bloom words and table probes hit L2.
#define SQ_PCHUNK 2048
#define SQ_PBLOCK 256
#define SQ_PUNROLL (SQ_PCHUNK / SQ_PBLOCK)

extern "C" __global__ void __launch_bounds__(SQ_PBLOCK) sq_joinprobe_kernel(SqIn in, i64 n, SqJoin jt, int keep_unmatched, int* __restrict__ slot_of,
                                                                             u32* __restrict__ chunk_counts, u32* __restrict__ err) {
  __shared__ u32 queue_s[SQ_PBLOCK / 32][SQ_PUNROLL * 32];
  __shared__ u64 queue_vs[SQ_PBLOCK / 32][SQ_PQMODE ? SQ_PUNROLL * 32 : 1];
  __shared__ u32 warp_total[SQ_PBLOCK / 32];
  bool any_err = false;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32* queue = queue_s[warp];
  u64* queue_v = queue_vs[warp];
  const u32 lanes_below = (1u << lane) - 1u;
  const u64 pol_keep = sq_l2_evict_last();
  const i64 n_chunks = (n + SQ_PCHUNK - 1) / SQ_PCHUNK;
  for (i64 chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const i64 base = chunk * SQ_PCHUNK + (i64)warp * (SQ_PUNROLL * 32);
    u32 out_rows = 0;  // per lane
    u32 queued = 0;    // warp-uniform
    // ---- phase A: streaming Filter + hash + Bloom test (see joinagg.cuh)
    u64 qv[SQ_PUNROLL];
    u32 bits[SQ_PUNROLL], bw[SQ_PUNROLL];
    bool live[SQ_PUNROLL], kept[SQ_PUNROLL];
#pragma unroll
    for (int u = 0; u < SQ_PUNROLL; u++) {
      const i64 r = base + u * 32 + lane;
      const bool inb = r < n;
      SqProbe p;
      bool e0 = false, e1 = false;
      sq_probe_row(in, inb ? r : n - 1, p, e0, e1);
      kept[u] = inb && p.pass;
      live[u] = kept[u];
#if SQ_JMATCH
      live[u] = live[u] && p.knull == 0u;  // SQL semantics: a NULL key never joins
#endif
      any_err |= (inb && e0) || (kept[u] && e1);
      qv[u] = SQ_PQMODE ? sq_probe_qv(p) : 0ULL;
      bits[u] = sq_bloom_bits(p.h);
      bw[u] = sq_bloom_word(p.h, jt.bloom_mask);
    }
#pragma unroll
    for (int u = 0; u < SQ_PUNROLL; u++) bw[u] = live[u] ? sq_ld_u32_l2(&jt.bloom[bw[u]], pol_keep) : 0u;
#pragma unroll
    for (int u = 0; u < SQ_PUNROLL; u++) {
      const i64 r = base + u * 32 + lane;
      const bool cand = live[u] && (bw[u] & bits[u]) == bits[u];
      const u32 m = __ballot_sync(0xffffffffu, cand);
      if (cand) {
        const u32 pos = queued + __popc(m & lanes_below);
        queue[pos] = (u32)(u * 32 + lane);
        if (SQ_PQMODE) queue_v[pos] = qv[u];
      } else if (r < n) {
        slot_of[r] = kept[u] ? -1 : -2;
        out_rows += (kept[u] && keep_unmatched) ? 1u : 0u;
      }
      queued += __popc(m);
    }
    __syncwarp();
    // ---- phase B: the candidates, compacted — exact probe with all lanes busy
    for (u32 i = lane; i < queued; i += 32) {
      const i64 r = base + queue[i];
      SqProbe p;
#if SQ_PQMODE
      sq_probe_unq(queue_v[i], p);
#else
      bool e0 = false, e1 = false;
      sq_probe_row(in, r, p, e0, e1);
#endif
      const int slot = sq_join_find(jt, p);
      slot_of[r] = slot;
      out_rows += slot >= 0 ? (jt.unique ? 1u : __ldg(&jt.slot_count[slot])) : (keep_unmatched ? 1u : 0u);
    }
    __syncwarp();
    // ---- chunk total
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) out_rows += __shfl_xor_sync(0xffffffffu, out_rows, d);
    if (lane == 0) warp_total[warp] = out_rows;
    __syncthreads();
    if (threadIdx.x == 0) {
      u32 t = 0;
#pragma unroll
      for (int w = 0; w < SQ_PBLOCK / 32; w++) t += warp_total[w];
      chunk_counts[chunk] = t;
    }
    __syncthreads();
  }
  if (any_err) atomicOr(err, 1u);
}
