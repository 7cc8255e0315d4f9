// sqlrs_b200 — hash join build / probe kernels for sm_100a (reference src/executor/join/hash_join.rs).
// Build (:161-187): instead of HashMap<u64, Vec<usize>> the distinct keys go into an open-addressed
// table (+ a blocked Bloom filter) and — only when some key repeats — the build row ids into a CSR
// array grouped by key, ascending per key.  Probe (:208-248): the fused JIT kernel csrc/jit/joinprobe.cuh
// finds the slot of every probe row; k_join_probe_emit here turns slots + per-chunk offsets into
// (build row, probe row) pairs exactly in the reference's order, with no per-row allocation.  All kernels are
// HBM-latency bound (random 8-byte reads of the table); grids are multiples of the SM count.
#include <cub/device/device_radix_sort.cuh>

#include "kernels_aot.hpp"

namespace sq {
namespace {

constexpr int kBlock = 256;
inline unsigned grid_for(int64_t items, int per_block, int64_t cap = 148 * 16) {
  int64_t g = div_up(items, per_block);
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (unsigned)g;
}

__device__ __forceinline__ uint32_t mix32(uint64_t h) {
  h ^= h >> 33;
  h *= 0xff51afd7ed558ccdULL;
  h ^= h >> 29;
  return (uint32_t)h;
}

// does build row `b` carry the identity (h, keys)?  keys/knull of the other side given by pointer+index
__device__ __forceinline__ bool same_key(const JoinTableView& t, int64_t b, uint64_t h, const uint64_t* okeys, int64_t ostride, int64_t oi) {
  if (t.h[b] != h) return false;
  if (!t.match_keys) return true;
  for (int k = 0; k < t.n_keys; k++)
    if (t.keys[(size_t)k * t.n_build + b] != okeys[(size_t)k * ostride + oi]) return false;
  return true;
}

// Every insert is a chain of dependent random accesses (slot -> CAS -> count); a thread therefore works on kInsertBatch
// rows at a time and issues each level of the chain for all of them before it waits (measured with one row at a
// time: 88 long-scoreboard stall cycles per issued instruction, 10 % issue utilisation).
constexpr int kInsertBatch = 4;
// Inserts do not count rows per key: they only flag (*has_dups) that some key occurred twice — the common build side is a
// primary key, for which neither per-slot counts nor the CSR row lists are ever needed (k_join_count runs otherwise).
__global__ void __launch_bounds__(kBlock) k_join_insert(JoinTableView t, int32_t* __restrict__ row_slot, uint32_t* __restrict__ has_dups) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const uint32_t mask = t.capacity - 1;
  bool dup = false;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < t.n_build; base += stride * kInsertBatch) {
    int64_t i[kInsertBatch];
    bool live[kInsertBatch];
    uint64_t h[kInsertBatch];
    uint32_t s[kInsertBatch];
    long long rep[kInsertBatch];
#pragma unroll
    for (int u = 0; u < kInsertBatch; u++) {
      i[u] = base + u * stride;
      const bool inb = i[u] < t.n_build;
      const int64_t r = inb ? i[u] : 0;
      const bool kept = !t.build_keep || ((t.build_keep[r >> 5] >> (r & 31)) & 1u);            // passes the fused Filter
      const bool null_key = t.match_keys && t.knull[r] != 0u;                                  // SQL semantics: a NULL key never joins
      h[u] = t.h[r];
      live[u] = inb && kept && !null_key;
      s[u] = mix32(h[u]) & mask;
    }
#pragma unroll
    for (int u = 0; u < kInsertBatch; u++) rep[u] = live[u] ? *((volatile long long*)&t.slot_rep[s[u]]) : -1;
#pragma unroll
    for (int u = 0; u < kInsertBatch; u++) {
      if (!live[u]) {
        if (i[u] < t.n_build) row_slot[i[u]] = -1;
        continue;
      }
      uint32_t slot = s[u];
      long long r = rep[u];
      for (;;) {
        if (r < 0) {
          const long long old = (long long)atomicCAS((unsigned long long*)&t.slot_rep[slot], (unsigned long long)-1LL, (unsigned long long)i[u]);
          r = old < 0 ? i[u] : old;
        }
        if (r == i[u]) break;
        if (same_key(t, r, h[u], t.keys, t.n_build, i[u])) {
          dup = true;
          break;
        }
        slot = (slot + 1) & mask;
        r = *((volatile long long*)&t.slot_rep[slot]);
      }
      row_slot[i[u]] = (int32_t)slot;
      if (t.bloom) atomicOr(&t.bloom[join_bloom_word(h[u], t.bloom_mask)], join_bloom_bits(h[u]));
    }
  }
  if (__any_sync(0xffffffffu, dup) && (threadIdx.x & 31) == 0) *has_dups = 1u;
}

// The kv layout (JoinTableView::kv): the key word of a slot is claimed by CAS against kJoinKvEmpty, the representative row is
// the smallest build row of the key (atomicMin), so the table is deterministic whatever the thread order.  One atomic on a
// fresh line per insert, no dependent read of a per-row key array on a collision.
__global__ void __launch_bounds__(kBlock) k_join_insert_kv(JoinTableView t, int32_t* __restrict__ row_slot, uint32_t* __restrict__ misc) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const uint32_t mask = t.capacity - 1;
  bool dup = false, sentinel = false, full = false;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < t.n_build; base += stride * kInsertBatch) {
    int64_t i[kInsertBatch];
    bool live[kInsertBatch];
    uint64_t h[kInsertBatch], key[kInsertBatch], seen[kInsertBatch];
    uint32_t s[kInsertBatch];
#pragma unroll
    for (int u = 0; u < kInsertBatch; u++) {
      i[u] = base + u * stride;
      const bool inb = i[u] < t.n_build;
      const int64_t r = inb ? i[u] : 0;
      const bool kept = !t.build_keep || ((t.build_keep[r >> 5] >> (r & 31)) & 1u);
      const bool null_key = t.knull[r] != 0u;  // SQL semantics: a NULL key never joins
      h[u] = t.h[r];
      key[u] = t.keys[r];
      live[u] = inb && kept && !null_key;
      if (live[u] && key[u] == kJoinKvEmpty) {
        sentinel = true;
        live[u] = false;
      }
      s[u] = mix32(h[u]) & mask;
    }
#pragma unroll
    for (int u = 0; u < kInsertBatch; u++) seen[u] = live[u] ? *((volatile unsigned long long*)&t.kv[2 * (size_t)s[u]]) : 0ULL;
#pragma unroll
    for (int u = 0; u < kInsertBatch; u++) {
      if (!live[u]) {
        if (i[u] < t.n_build) row_slot[i[u]] = -1;
        continue;
      }
      uint32_t slot = s[u];
      unsigned long long cur = seen[u];
      bool placed = true;
      for (uint32_t probes = 0;; probes++) {
        if (cur == kJoinKvEmpty) cur = atomicCAS((unsigned long long*)&t.kv[2 * (size_t)slot], (unsigned long long)kJoinKvEmpty, (unsigned long long)key[u]);
        if (cur == kJoinKvEmpty) break;  // claimed
        if (cur == key[u]) {
          dup = true;
          break;
        }
        if (probes >= mask) {  // table full (sized from a stale hint): flagged, the host rebuilds
          full = true;
          placed = false;
          break;
        }
        slot = (slot + 1) & mask;
        cur = *((volatile unsigned long long*)&t.kv[2 * (size_t)slot]);
      }
      if (!placed) {
        row_slot[i[u]] = -1;
        continue;
      }
      atomicMin((unsigned long long*)&t.kv[2 * (size_t)slot + 1], (unsigned long long)i[u]);
      row_slot[i[u]] = (int32_t)slot;
      if (t.bloom) atomicOr(&t.bloom[join_bloom_word(h[u], t.bloom_mask)], join_bloom_bits(h[u]));
    }
  }
  if (__any_sync(0xffffffffu, dup) && (threadIdx.x & 31) == 0) misc[0] = 1u;
  if (__any_sync(0xffffffffu, sentinel) && (threadIdx.x & 31) == 0) misc[4] = 1u;
  if (__any_sync(0xffffffffu, full) && (threadIdx.x & 31) == 0) misc[5] = 1u;
}

// only when some key repeats: rows per slot and the largest such count
__global__ void __launch_bounds__(kBlock) k_join_count(JoinTableView t, const int32_t* __restrict__ row_slot, uint32_t* __restrict__ max_count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  uint32_t local_max = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t.n_build; i += stride) {
    const int32_t s = row_slot[i];
    if (s < 0) continue;
    const uint32_t c = atomicAdd(&t.slot_count[s], 1u) + 1u;
    local_max = c > local_max ? c : local_max;
  }
  for (int d = 16; d > 0; d >>= 1) {
    const uint32_t o = __shfl_xor_sync(0xffffffffu, local_max, d);
    local_max = o > local_max ? o : local_max;
  }
  if ((threadIdx.x & 31) == 0 && local_max) atomicMax(max_count, local_max);
}

__global__ void __launch_bounds__(kBlock) k_join_fill(JoinTableView t, const int32_t* __restrict__ row_slot, uint32_t* __restrict__ slot_fill) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t.n_build; i += stride) {
    const int32_t s = row_slot[i];
    if (s < 0) continue;
    const uint64_t pos = t.slot_start[s] + atomicAdd(&slot_fill[s], 1u);
    t.rows[pos] = i;
  }
}

// ranges are short (bounded by the caller): insertion sort by one thread
__global__ void __launch_bounds__(kBlock) k_join_sort_ranges(JoinTableView t) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < t.capacity; s += stride) {
    const uint32_t c = t.slot_count[s];
    if (c < 2) continue;
    int64_t* r = t.rows + t.slot_start[s];
    for (uint32_t a = 1; a < c; a++) {
      const int64_t v = r[a];
      uint32_t b = a;
      while (b > 0 && r[b - 1] > v) {
        r[b] = r[b - 1];
        b--;
      }
      r[b] = v;
    }
  }
}

// one CTA per 2048-row chunk, 8 consecutive probe rows per thread: CTA-wide exclusive prefix of the per-row output
// counts, then the pairs — positions ascend with the probe row, per probe row with the build insertion order
__global__ void __launch_bounds__(kBlock) k_join_probe_emit(JoinTableView t, const int32_t* __restrict__ slot_of,
                                                             const unsigned long long* __restrict__ chunk_offsets, int64_t n_probe, int keep_unmatched,
                                                             int64_t* __restrict__ li, uint32_t* __restrict__ ri) {
  constexpr int kPer = kProbeChunk / kBlock;  // 8
  __shared__ uint32_t warp_sums[kBlock / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t n_chunks = (n_probe + kProbeChunk - 1) / kProbeChunk;
  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const int64_t r0 = chunk * kProbeChunk + (int64_t)threadIdx.x * kPer;
    int32_t slot[kPer];
    uint32_t cnt[kPer];
    uint32_t local = 0;
#pragma unroll
    for (int j = 0; j < kPer; j++) {
      slot[j] = r0 + j < n_probe ? slot_of[r0 + j] : -2;
      cnt[j] = slot[j] >= 0 ? (t.unique ? 1u : t.slot_count[slot[j]]) : ((slot[j] == -1 && keep_unmatched) ? 1u : 0u);
      local += cnt[j];
    }
    uint32_t x = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    unsigned long long o = chunk_offsets[chunk] + (x - local);
    for (int w = 0; w < wid; w++) o += warp_sums[w];
#pragma unroll
    for (int j = 0; j < kPer; j++) {
      if (slot[j] >= 0) {
        if (t.unique) {
          li[o] = t.slot_rep[(size_t)slot[j] * t.rep_stride];
          ri[o] = (uint32_t)(r0 + j);
        } else {
          const int64_t* src = t.rows + t.slot_start[slot[j]];
          for (uint32_t q = 0; q < cnt[j]; q++) {
            li[o + q] = src[q];
            ri[o + q] = (uint32_t)(r0 + j);
          }
        }
      } else if (cnt[j]) {  // Right/Full: (NULL, row), hash_join.rs:242-246
        li[o] = -1;
        ri[o] = (uint32_t)(r0 + j);
      }
      o += cnt[j];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kBlock) k_slot_keep_bitmap(const int32_t* __restrict__ slot_of, int64_t n, uint32_t* __restrict__ bitmap) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_up = (n + 31) & ~(int64_t)31;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_up; r += stride) {
    const uint32_t w = __ballot_sync(0xffffffffu, r < n && slot_of[r] != -2);
    if (lane == 0) bitmap[r >> 5] = w;
  }
}

template <typename Idx>
__global__ void __launch_bounds__(kBlock) k_mark_bits(const Idx* __restrict__ idx, int64_t m, uint32_t* __restrict__ bitmap) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += stride) {
    const long long i = (long long)idx[k];
    if (i < 0) continue;
    atomicOr(&bitmap[i >> 5], 1u << (i & 31));
  }
}

__global__ void __launch_bounds__(kBlock) k_bitmap_not(const uint32_t* __restrict__ src, int64_t n, uint32_t* __restrict__ dst,
                                                        const uint32_t* __restrict__ and_mask) {
  const int64_t words = (n + 31) >> 5;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += stride) {
    uint32_t v = ~src[w];
    if (and_mask) v &= and_mask[w];
    if (w == words - 1 && (n & 31)) v &= (1u << (n & 31)) - 1u;
    dst[w] = v;
  }
}

// ---- large exclusive scan: 4096-entry chunks -> chunk sums -> scan of sums -> apply
constexpr int kScanChunk = 4096;

__global__ void __launch_bounds__(kBlock) k_scan_chunk_sums(const uint32_t* __restrict__ counts, int64_t m, unsigned long long* __restrict__ sums) {
  __shared__ unsigned long long warp_sums[kBlock / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanChunk;
  unsigned long long local = 0;
  for (int j = threadIdx.x; j < kScanChunk; j += kBlock)
    if (base + j < m) local += counts[base + j];
  for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int w = 0; w < kBlock / 32; w++) s += warp_sums[w];
    sums[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(1024) k_scan_u64_inplace(unsigned long long* __restrict__ v, int64_t m, unsigned long long* __restrict__ total) {
  __shared__ unsigned long long warp_sums[32];
  __shared__ unsigned long long carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < m; base += 1024) {
    const int64_t i = base + tid;
    const unsigned long long x0 = i < m ? v[i] : 0;
    unsigned long long x = x0;
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      unsigned long long s = warp_sums[lane];
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, s, d);
        if (lane >= d) s += y;
      }
      warp_sums[lane] = s;
    }
    __syncthreads();
    const unsigned long long carry = carry_s;
    if (i < m) v[i] = carry + (wid ? warp_sums[wid - 1] : 0) + (x - x0);
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_sums[31];
    __syncthreads();
  }
  if (tid == 0) *total = carry_s;
}

// one CTA per chunk: 256 threads x 16 consecutive entries
__global__ void __launch_bounds__(kBlock) k_scan_chunk_apply(const uint32_t* __restrict__ counts, int64_t m, const unsigned long long* __restrict__ chunk_base,
                                                              unsigned long long* __restrict__ offsets) {
  __shared__ unsigned long long warp_sums[kBlock / 32];
  const int per = kScanChunk / kBlock;  // 16
  const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * per;
  unsigned long long vals[per];
  unsigned long long local = 0;
#pragma unroll
  for (int j = 0; j < per; j++) {
    vals[j] = (base + j < m) ? counts[base + j] : 0;
    local += vals[j];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned long long x = local;
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  unsigned long long before = chunk_base[blockIdx.x] + (x - local);
  for (int w = 0; w < wid; w++) before += warp_sums[w];
#pragma unroll
  for (int j = 0; j < per; j++) {
    if (base + j < m) offsets[base + j] = before;
    before += vals[j];
  }
}

__global__ void __launch_bounds__(kBlock) k_iota_i64(int64_t* __restrict__ dst, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = i;
}
__global__ void __launch_bounds__(kBlock) k_slot_keys(const int32_t* __restrict__ row_slot, int64_t n, uint32_t capacity, uint32_t* __restrict__ keys) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    keys[i] = row_slot[i] < 0 ? capacity : (uint32_t)row_slot[i];  // NULL-key rows sort past every slot
}

#define SQ_LAUNCH_CHECK()   \
  do {                      \
    count_launch();         \
    SQ_CUDA(cudaGetLastError()); \
  } while (0)

}  // namespace

uint32_t join_bloom_words(int64_t n_build) {
  static const int shift = [] {  // tuning experiments: SQLRS_B200_BLOOM_SHIFT (words >= n_build >> shift; 1 = 16+ bits per key)
    const char* e = std::getenv("SQLRS_B200_BLOOM_SHIFT");
    return e ? std::min(6, std::max(0, atoi(e))) : 1;
  }();
  uint64_t w = 1024;
  while (w < ((uint64_t)n_build >> shift) && w < (1ULL << 25)) w <<= 1;
  return (uint32_t)w;
}

size_t scan_scratch_entries(int64_t m) { return (size_t)div_up(m, kScanChunk) + 1; }

void launch_scan_u32_large(const uint32_t* counts, int64_t m, unsigned long long* offsets, unsigned long long* total,
                           unsigned long long* scratch, cudaStream_t stream) {
  if (m <= 0) {
    SQ_CUDA(cudaMemsetAsync(total, 0, 8, stream));
    return;
  }
  const int64_t chunks = div_up(m, kScanChunk);
  k_scan_chunk_sums<<<(unsigned)chunks, kBlock, 0, stream>>>(counts, m, scratch);
  SQ_LAUNCH_CHECK();
  k_scan_u64_inplace<<<1, 1024, 0, stream>>>(scratch, chunks, total);
  SQ_LAUNCH_CHECK();
  k_scan_chunk_apply<<<(unsigned)chunks, kBlock, 0, stream>>>(counts, m, scratch, offsets);
  SQ_LAUNCH_CHECK();
}

void launch_join_insert(const JoinTableView& t, int32_t* row_slot, uint32_t* has_dups, cudaStream_t stream) {
  if (t.n_build <= 0) return;
  k_join_insert<<<grid_for(div_up(t.n_build, kInsertBatch), kBlock, 148 * 32), kBlock, 0, stream>>>(t, row_slot, has_dups);
  SQ_LAUNCH_CHECK();
}
void launch_join_insert_kv(const JoinTableView& t, int32_t* row_slot, uint32_t* misc, cudaStream_t stream) {
  if (t.n_build <= 0) return;
  k_join_insert_kv<<<grid_for(div_up(t.n_build, kInsertBatch), kBlock, 148 * 32), kBlock, 0, stream>>>(t, row_slot, misc);
  SQ_LAUNCH_CHECK();
}
void launch_join_count(const JoinTableView& t, const int32_t* row_slot, uint32_t* max_count, cudaStream_t stream) {
  if (t.n_build <= 0) return;
  k_join_count<<<grid_for(t.n_build, kBlock), kBlock, 0, stream>>>(t, row_slot, max_count);
  SQ_LAUNCH_CHECK();
}
void launch_join_fill(const JoinTableView& t, const int32_t* row_slot, uint32_t* slot_fill, cudaStream_t stream) {
  if (t.n_build <= 0) return;
  k_join_fill<<<grid_for(t.n_build, kBlock), kBlock, 0, stream>>>(t, row_slot, slot_fill);
  SQ_LAUNCH_CHECK();
}
void launch_join_sort_ranges(const JoinTableView& t, cudaStream_t stream) {
  k_join_sort_ranges<<<grid_for(t.capacity, kBlock), kBlock, 0, stream>>>(t);
  SQ_LAUNCH_CHECK();
}

void join_fill_sorted(const JoinTableView& t, const int32_t* row_slot, cudaStream_t stream) {
  // stable LSD radix sort of (slot, row id): equal slots keep ascending row ids
  const int64_t n = t.n_build;
  if (n <= 0) return;
  uint32_t *k_in = nullptr, *k_out = nullptr;
  int64_t *v_in = nullptr, *v_out = nullptr;
  k_in = (decltype(k_in))scratch_alloc((size_t)n * 4, stream);
  k_out = (decltype(k_out))scratch_alloc((size_t)n * 4, stream);
  v_in = (decltype(v_in))scratch_alloc((size_t)n * 8, stream);
  v_out = (decltype(v_out))scratch_alloc((size_t)n * 8, stream);
  k_slot_keys<<<grid_for(n, kBlock), kBlock, 0, stream>>>(row_slot, n, t.capacity, k_in);
  SQ_LAUNCH_CHECK();
  k_iota_i64<<<grid_for(n, kBlock), kBlock, 0, stream>>>(v_in, n);
  SQ_LAUNCH_CHECK();
  int bits = 1;
  while ((1ull << bits) <= t.capacity) bits++;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, v_out, (int)n, 0, bits, stream);
  void* tmp = nullptr;
  tmp = (decltype(tmp))scratch_alloc(tmp_bytes ? tmp_bytes : 16, stream);
  SQ_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, (int)n, 0, bits, stream));
  count_launch(4);
  // rows of NULL-key build rows (slot = capacity) sort last and are never referenced by a slot range
  SQ_CUDA(cudaMemcpyAsync(t.rows, v_out, (size_t)n * 8, cudaMemcpyDeviceToDevice, stream));
  scratch_free(tmp, stream);
  scratch_free(k_in, stream);
  scratch_free(k_out, stream);
  scratch_free(v_in, stream);
  scratch_free(v_out, stream);
}

void launch_join_probe_emit(const JoinTableView& t, const int32_t* slot_of, const unsigned long long* chunk_offsets, int64_t n_probe,
                            int keep_unmatched, int64_t* li, uint32_t* ri, cudaStream_t stream) {
  if (n_probe <= 0) return;
  k_join_probe_emit<<<grid_for(div_up(n_probe, kProbeChunk), 1, 148 * 8), kBlock, 0, stream>>>(t, slot_of, chunk_offsets, n_probe, keep_unmatched, li, ri);
  SQ_LAUNCH_CHECK();
}
void launch_slot_keep_bitmap(const int32_t* slot_of, int64_t n, uint32_t* bitmap, cudaStream_t stream) {
  if (n <= 0) return;
  k_slot_keep_bitmap<<<grid_for(n, kBlock), kBlock, 0, stream>>>(slot_of, n, bitmap);
  SQ_LAUNCH_CHECK();
}
void launch_mark_bits_i64(const int64_t* idx, int64_t m, uint32_t* bitmap, cudaStream_t stream) {
  if (m <= 0) return;
  k_mark_bits<int64_t><<<grid_for(m, kBlock), kBlock, 0, stream>>>(idx, m, bitmap);
  SQ_LAUNCH_CHECK();
}
void launch_mark_bits_u32(const uint32_t* idx, int64_t m, uint32_t* bitmap, cudaStream_t stream) {
  if (m <= 0) return;
  k_mark_bits<uint32_t><<<grid_for(m, kBlock), kBlock, 0, stream>>>(idx, m, bitmap);
  SQ_LAUNCH_CHECK();
}
void launch_bitmap_not(const uint32_t* src, int64_t n, uint32_t* dst, cudaStream_t stream, const uint32_t* and_mask) {
  if (n <= 0) return;
  k_bitmap_not<<<grid_for(div_up(n, 32), kBlock), kBlock, 0, stream>>>(src, n, dst, and_mask);
  SQ_LAUNCH_CHECK();
}

}  // namespace sq
