// sqlrs_b200 — ahead-of-time compiled CUDA kernels for sm_100a: synthetic table generator,
// bitmap utilities, warp-ballot stream compaction, gathers (arrow `take`), group-table maintenance.
// All of it is HBM-bound integer / byte work: coalesced 4/8-byte lanes, whole-word bitmap stores
// through __ballot_sync, grids sized in multiples of the SM count.
#include <cub/device/device_radix_sort.cuh>

#include "kernels_aot.hpp"

#include "../../include/sqlrs_tpch_spec.h"

namespace sq {

namespace {

constexpr int kBlock = 256;

inline unsigned grid_for(int64_t items, int per_block, int64_t cap = 148 * 32) {
  int64_t g = div_up(items, per_block);
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (unsigned)g;
}

// ------------------------------------------------------------------ generator
__global__ void __launch_bounds__(kBlock) k_tpch_generate(int table, int col, int64_t row_begin, int64_t n, int64_t n_customer,
                                                           int flags_mode, uint64_t* __restrict__ dst) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = sqlrs_tpch_cell(table, col, row_begin + i, n_customer, flags_mode);
}

// ------------------------------------------------------------------ bitmaps
__global__ void __launch_bounds__(kBlock) k_count_bits(const uint32_t* __restrict__ bm, int64_t n, unsigned long long* out) {
  const int64_t words = (n + 31) >> 5;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long local = 0;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += stride) {
    uint32_t v = bm[w];
    if (w == words - 1 && (n & 31)) v &= (1u << (n & 31)) - 1u;
    local += __popc(v);
  }
  for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

// one thread per destination word; partial first/last words are OR-ed in atomically
__global__ void __launch_bounds__(kBlock) k_bitmap_append(uint32_t* __restrict__ dst, int64_t dst_off, const uint32_t* __restrict__ src,
                                                           int64_t n) {
  const int64_t first_word = dst_off >> 5, last_word = (dst_off + n - 1) >> 5;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t src_words = (n + 31) >> 5;
  for (int64_t w = first_word + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w <= last_word; w += stride) {
    // destination word w covers destination bits [w*32, w*32+32) = source bits [w*32-dst_off, ...)
    const int64_t s0 = w * 32 - dst_off;  // may be negative for the first word
    uint32_t v = 0;
    if (!src) {
      v = 0xffffffffu;
    } else {
      const int64_t sw = s0 >> 5;  // floor
      const int sh = (int)(s0 & 31);
      const uint32_t lo = (sw >= 0 && sw < src_words) ? src[sw] : 0u;
      const uint32_t hi = (sw + 1 >= 0 && sw + 1 < src_words) ? src[sw + 1] : 0u;
      v = sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
    }
    // mask to the valid source range [0, n)
    uint32_t mask = 0xffffffffu;
    if (s0 < 0) mask &= 0xffffffffu << (int)(-s0);
    const int64_t end = s0 + 32 - n;  // bits past the end
    if (end > 0) mask &= end >= 32 ? 0u : (0xffffffffu >> (int)end);
    v &= mask;
    if (mask == 0xffffffffu) dst[w] = v;
    else if (v) atomicOr(&dst[w], v);
  }
}

// ------------------------------------------------------------------ compaction
constexpr int kChunkWords = 64;  // 2048 rows per warp-chunk

__global__ void __launch_bounds__(kBlock) k_compact_count(const uint32_t* __restrict__ keep, int64_t n_words, int64_t n_chunks,
                                                           uint32_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kBlock / 32);
  for (int64_t c = warp; c < n_chunks; c += nwarps) {
    const int64_t w0 = c * kChunkWords;
    uint32_t cnt = 0;
    for (int j = lane; j < kChunkWords; j += 32)
      if (w0 + j < n_words) cnt += __popc(keep[w0 + j]);
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    if (lane == 0) counts[c] = cnt;
  }
}

// exclusive scan of m u32 counts into u64 offsets; one CTA walks the array with a running carry
__global__ void __launch_bounds__(1024) k_scan_u32(const uint32_t* __restrict__ counts, int64_t m, unsigned long long* __restrict__ offsets,
                                                    unsigned long long* __restrict__ total) {
  __shared__ unsigned long long warp_sums[32];
  __shared__ unsigned long long carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < m; base += 1024) {
    const int64_t i = base + tid;
    const unsigned long long v = i < m ? counts[i] : 0;
    unsigned long long x = v;
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
      warp_sums[lane] = s;  // inclusive
    }
    __syncthreads();
    const unsigned long long carry = carry_s;
    const unsigned long long before = carry + (wid ? warp_sums[wid - 1] : 0) + (x - v);
    if (i < m) offsets[i] = before;
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_sums[31];
    __syncthreads();
  }
  if (tid == 0) *total = carry_s;
}

__global__ void __launch_bounds__(kBlock) k_compact_write(const uint32_t* __restrict__ keep, int64_t n_words, int64_t n_chunks,
                                                           const unsigned long long* __restrict__ offsets, uint32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (kBlock / 32);
  for (int64_t c = warp; c < n_chunks; c += nwarps) {
    unsigned long long off = offsets[c];
    const int64_t w0 = c * kChunkWords;
    // lanes fetch two words each, then the warp walks the 64 words with the bits spread over lanes
    const uint32_t mine0 = (w0 + lane < n_words) ? keep[w0 + lane] : 0u;
    const uint32_t mine1 = (w0 + 32 + lane < n_words) ? keep[w0 + 32 + lane] : 0u;
    for (int j = 0; j < kChunkWords; j++) {
      const uint32_t word = __shfl_sync(0xffffffffu, j < 32 ? mine0 : mine1, j & 31);
      if (word == 0u) continue;
      if ((word >> lane) & 1u) out[off + __popc(word & ((1u << lane) - 1u))] = (uint32_t)((w0 + j) * 32 + lane);
      off += __popc(word);
    }
  }
}

// ------------------------------------------------------------------ gather
template <typename Idx>
__device__ __forceinline__ bool idx_ok(Idx i);
template <>
__device__ __forceinline__ bool idx_ok<uint32_t>(uint32_t) { return true; }
template <>
__device__ __forceinline__ bool idx_ok<int64_t>(int64_t i) { return i >= 0; }

// WIDTH: 8 / 4 bytes, 0 = bit-packed Boolean values
template <int WIDTH, typename Idx>
__global__ void __launch_bounds__(kBlock) k_gather(const void* __restrict__ src, const uint32_t* __restrict__ src_valid,
                                                    const Idx* __restrict__ idx, int64_t m, void* __restrict__ dst,
                                                    uint32_t* __restrict__ dst_valid) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t m_up = (m + 31) & ~(int64_t)31;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m_up; k += stride) {
    const bool inb = k < m;
    Idx i = inb ? idx[k] : (Idx)0;
    const bool ok = inb && idx_ok<Idx>(i);
    const uint64_t r = ok ? (uint64_t)i : 0;
    bool valid = ok;
    if (ok && src_valid) valid = (src_valid[r >> 5] >> (r & 31)) & 1u;
    if (WIDTH == 8) {
      if (inb) ((uint64_t*)dst)[k] = ok ? ((const uint64_t*)src)[r] : 0ULL;
    } else if (WIDTH == 4) {
      if (inb) ((uint32_t*)dst)[k] = ok ? ((const uint32_t*)src)[r] : 0u;
    } else {
      const bool bit = ok && ((((const uint32_t*)src)[r >> 5] >> (r & 31)) & 1u);
      const uint32_t w = __ballot_sync(0xffffffffu, bit);
      if (lane == 0) ((uint32_t*)dst)[k >> 5] = w;
    }
    if (dst_valid) {
      const uint32_t w = __ballot_sync(0xffffffffu, valid);
      if (lane == 0) dst_valid[k >> 5] = w;
    }
  }
}

template <typename Idx>
void launch_gather_t(int width, const void* src, const uint32_t* src_valid, const Idx* idx, int64_t m, void* dst, uint32_t* dst_valid,
                     cudaStream_t stream) {
  if (m <= 0) return;
  unsigned grid = grid_for(m, kBlock);
  if (width == 8) k_gather<8, Idx><<<grid, kBlock, 0, stream>>>(src, src_valid, idx, m, dst, dst_valid);
  else if (width == 4) k_gather<4, Idx><<<grid, kBlock, 0, stream>>>(src, src_valid, idx, m, dst, dst_valid);
  else k_gather<0, Idx><<<grid, kBlock, 0, stream>>>(src, src_valid, idx, m, dst, dst_valid);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(kBlock) k_fill_u64(uint64_t* __restrict__ dst, int64_t n, uint64_t value) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = value;
}

__global__ void __launch_bounds__(kBlock) k_idx_valid(const int64_t* __restrict__ idx, int64_t m, uint32_t* __restrict__ valid_out) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t m_up = (m + 31) & ~(int64_t)31;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m_up; k += stride) {
    const uint32_t w = __ballot_sync(0xffffffffu, k < m && idx[k] >= 0);
    if (lane == 0) valid_out[k >> 5] = w;
  }
}

// ------------------------------------------------------------------ group table maintenance
// Packed (row-major) form of a group table, used for finalisation and for the partial/final exchange:
// row 0 = header {number of groups, words per row, 0...}; row 1+i = [hash, min_row, knull, key bits x K,
// accumulator words x W].  Rows land in arbitrary order (the consumer orders by min_row).
__global__ void __launch_bounds__(kBlock) k_table_pack(TableView t, int n_keys, int n_acc, uint64_t* __restrict__ dst, unsigned long long cap_rows,
                                                        int mark_unchecked) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const int words = 3 + n_keys + n_acc;
  // the table was filled by a launch nobody has checked yet: if that launch ran out of per-CTA slots (status bit 0) or flagged an
  // arithmetic error, the header count is pushed beyond any capacity — every consumer of the buffer then takes its overflow path
  if (mark_unchecked && blockIdx.x == 0 && threadIdx.x == 0 && ((t.counters[2] & 1u) || (t.counters[3] & 1u)))
    atomicAdd((unsigned long long*)dst, 1ULL << 40);
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < t.capacity; s += stride) {
    if (t.state[s] != 2u) continue;
    const unsigned long long o = atomicAdd((unsigned long long*)dst, 1ULL);
    if (o >= cap_rows) continue;  // header keeps counting: the consumer sees count > capacity
    uint64_t* row = dst + (size_t)(1 + o) * words;
    row[0] = t.hash[s];
    row[1] = t.min_row[s];
    row[2] = t.knull[s];
    for (int k = 0; k < n_keys; k++) row[3 + k] = t.keys[(size_t)k * t.capacity + s];
    for (int w = 0; w < n_acc; w++) row[3 + n_keys + w] = t.acc[(size_t)w * t.capacity + s];
  }
}

// radix partitioning for the multi-GPU exchange: group rows packed into n_parts regions by identity hash mod n_parts;
// region q starts at dst + q * (cap_rows + 1) * words, its row 0 is the header {count}
__global__ void __launch_bounds__(kBlock) k_table_pack_partitioned(TableView t, int n_keys, int n_acc, uint64_t* __restrict__ dst, unsigned n_parts,
                                                                    unsigned long long cap_rows) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const int words = 3 + n_keys + n_acc;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < t.capacity; s += stride) {
    if (t.state[s] != 2u) continue;
    const uint64_t h = t.hash[s];
    uint64_t* region = dst + (size_t)(h % n_parts) * (cap_rows + 1) * words;
    const unsigned long long o = atomicAdd((unsigned long long*)region, 1ULL);
    if (o >= cap_rows) continue;  // header keeps counting: the consumer sees count > capacity
    uint64_t* row = region + (size_t)(1 + o) * words;
    row[0] = h;
    row[1] = t.min_row[s];
    row[2] = t.knull[s];
    for (int k = 0; k < n_keys; k++) row[3 + k] = t.keys[(size_t)k * t.capacity + s];
    for (int w = 0; w < n_acc; w++) row[3 + n_keys + w] = t.acc[(size_t)w * t.capacity + s];
  }
}

// occupied slots and their first-row ids, in table order (input of the ordering sort)
__global__ void __launch_bounds__(kBlock) k_table_list(TableView t, uint64_t* __restrict__ min_rows, uint32_t* __restrict__ slots,
                                                        uint32_t max_out, uint32_t* __restrict__ count) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < t.capacity; s += stride) {
    if (t.state[s] != 2u) continue;
    const uint32_t o = atomicAdd(count, 1u);
    if (o >= max_out) continue;
    min_rows[o] = t.min_row[s];
    slots[o] = s;
  }
}

// the same from a complete list of the occupied slots (tables filled by one fused probe+aggregate kernel keep one):
// n entries instead of a scan over the whole capacity
__global__ void __launch_bounds__(kBlock) k_table_list_from_slots(TableView t, const uint32_t* __restrict__ slot_list, uint32_t n,
                                                                   uint64_t* __restrict__ min_rows, uint32_t* __restrict__ slots) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t s = slot_list[i];
    min_rows[i] = t.min_row[s];
    slots[i] = s;
  }
}

// packed rows in the given slot order (first-appearance order after the sort); header written by thread 0
__global__ void __launch_bounds__(kBlock) k_table_pack_ordered(TableView t, int n_keys, int n_acc, const uint32_t* __restrict__ slots, uint32_t n,
                                                                uint64_t* __restrict__ dst) {
  const int words = 3 + n_keys + n_acc;
  const uint32_t stride = gridDim.x * blockDim.x;
  if (blockIdx.x == 0 && threadIdx.x == 0) dst[0] = n;
  for (uint32_t o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += stride) {
    const uint32_t s = slots[o];
    uint64_t* row = dst + (size_t)(1 + o) * words;
    row[0] = t.hash[s];
    row[1] = t.min_row[s];
    row[2] = t.knull[s];
    for (int k = 0; k < n_keys; k++) row[3 + k] = t.keys[(size_t)k * t.capacity + s];
    for (int w = 0; w < n_acc; w++) row[3 + n_keys + w] = t.acc[(size_t)w * t.capacity + s];
  }
}

__device__ __forceinline__ uint32_t mix32(uint64_t h) {  // same spreader as sq_mix32 (csrc/jit/prelude.cuh)
  h ^= h >> 33;
  h *= 0xff51afd7ed558ccdULL;
  h ^= h >> 29;
  return (uint32_t)h;
}

// every source slot is a distinct group already: claim the first free slot on its probe path
__global__ void __launch_bounds__(kBlock) k_table_rehash(TableView from, TableView to, int n_keys, int n_acc) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t mask = to.capacity - 1;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < from.capacity; s += stride) {
    if (from.state[s] != 2u) continue;
    const uint64_t h = from.hash[s];
    uint32_t d = mix32(h) & mask;
    while (atomicCAS(&to.state[d], 0u, 2u) != 0u) d = (d + 1) & mask;
    to.hash[d] = h;
    to.min_row[d] = from.min_row[s];
    to.knull[d] = from.knull[s];
    for (int k = 0; k < n_keys; k++) to.keys[(size_t)k * to.capacity + d] = from.keys[(size_t)k * from.capacity + s];
    for (int w = 0; w < n_acc; w++) to.acc[(size_t)w * to.capacity + d] = from.acc[(size_t)w * from.capacity + s];
  }
}

// partial -> final merge (multi-GPU group-by; also batches of host partials): n_bufs packed
// buffers of (cap_rows + 1) rows each; every row is a distinct group of its producer — find-or-insert it
// and fold its accumulator words in with the word's operator.  ops[w]: 0 add u64, 1 add f64, 2 min i64, 3 max i64.
__global__ void __launch_bounds__(kBlock) k_table_merge_packed(TableView t, int n_keys, int n_acc, const int* __restrict__ ops, int match_keys,
                                                                const uint64_t* __restrict__ src, int n_bufs, unsigned long long cap_rows) {
  const int words = 3 + n_keys + n_acc;
  const uint32_t mask = t.capacity - 1;
  const int64_t per_buf = (int64_t)cap_rows;
  const int64_t total = per_buf * n_bufs;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += stride) {
    const int b = (int)(j / per_buf);
    const int64_t i = j % per_buf;
    const uint64_t* buf = src + (size_t)b * (size_t)(cap_rows + 1) * words;
    if ((unsigned long long)i >= buf[0]) continue;
    const uint64_t* row = buf + (size_t)(1 + i) * words;
    const uint64_t h = row[0];
    const uint32_t kn = (uint32_t)row[2];
    uint32_t s = mix32(h) & mask;
    int slot = -1;
    for (uint32_t probes = 0; probes <= mask;) {
      const uint32_t st = *((volatile uint32_t*)&t.state[s]);
      if (st == 0u) {
        if (atomicCAS(&t.state[s], 0u, 1u) == 0u) {
          t.hash[s] = h;
          for (int k = 0; k < n_keys; k++) t.keys[(size_t)k * t.capacity + s] = row[3 + k];
          t.knull[s] = kn;
          __threadfence();
          atomicExch(&t.state[s], 2u);
          atomicAdd(&t.counters[0], 1u);
          slot = (int)s;
          break;
        }
        continue;
      }
      if (st == 1u) continue;
      __threadfence();
      if (*((volatile uint64_t*)&t.hash[s]) == h) {
        bool same = true;
        if (match_keys) {
          same = *((volatile uint32_t*)&t.knull[s]) == kn;
          for (int k = 0; same && k < n_keys; k++) same = *((volatile uint64_t*)&t.keys[(size_t)k * t.capacity + s]) == row[3 + k];
        }
        if (same) {
          slot = (int)s;
          break;
        }
      }
      s = (s + 1) & mask;
      probes++;
    }
    if (slot < 0) {
      atomicOr(&t.counters[2], 2u);
      continue;
    }
    // first-appearance order: the smaller global row id wins; with hash-only identity its keys win too — written by
    // k_table_merge_fixkeys once every min_row is final (one writer per slot, no interleaving of two key tuples)
    atomicMin((unsigned long long*)&t.min_row[slot], (unsigned long long)row[1]);
    for (int w = 0; w < n_acc; w++) {
      const uint64_t x = row[3 + n_keys + w];
      uint64_t* p = &t.acc[(size_t)w * t.capacity + slot];
      switch (ops[w]) {
        case 0: if (x) atomicAdd((unsigned long long*)p, (unsigned long long)x); break;
        case 1: atomicAdd((double*)p, __longlong_as_double((long long)x)); break;
        case 2: atomicMin((long long*)p, (long long)x); break;
        case 3: atomicMax((long long*)p, (long long)x); break;
        case 4: {  // COUNT under the overwrite quirk: (batch epoch << 40 | count) — the later batch wins, equal epochs add
          unsigned long long old = *p;
          for (;;) {
            const unsigned long long eo = old >> 40, ex = x >> 40;
            const unsigned long long nv = eo == ex ? old + (x & ((1ULL << 40) - 1)) : (ex > eo ? x : old);
            if (nv == old) break;
            const unsigned long long prev = atomicCAS((unsigned long long*)p, old, nv);
            if (prev == old) break;
            old = prev;
          }
          break;
        }
      }
    }
  }
}

// hash-only identity (reference quirk K2): after the merge, the partial group that supplied a slot's final min_row
// rewrites the slot's key tuple.  Global row ids are unique, so at most one source row matches per slot.
__global__ void __launch_bounds__(kBlock) k_table_merge_fixkeys(TableView t, int n_keys, int n_acc, const uint64_t* __restrict__ src, int n_bufs,
                                                                 unsigned long long cap_rows) {
  const int words = 3 + n_keys + n_acc;
  const uint32_t mask = t.capacity - 1;
  const int64_t per_buf = (int64_t)cap_rows;
  const int64_t total = per_buf * n_bufs;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += stride) {
    const int b = (int)(j / per_buf);
    const int64_t i = j % per_buf;
    const uint64_t* buf = src + (size_t)b * (size_t)(cap_rows + 1) * words;
    if ((unsigned long long)i >= buf[0]) continue;
    const uint64_t* row = buf + (size_t)(1 + i) * words;
    const uint64_t h = row[0];
    uint32_t s = mix32(h) & mask;
    for (uint32_t probes = 0; probes <= mask; probes++, s = (s + 1) & mask) {
      if (t.state[s] == 0u) break;
      if (t.hash[s] != h) continue;
      if (t.min_row[s] == row[1]) {
        for (int k = 0; k < n_keys; k++) t.keys[(size_t)k * t.capacity + s] = row[3 + k];
        t.knull[s] = (uint32_t)row[2];
      }
      break;
    }
  }
}

}  // namespace

void launch_table_merge_packed(const TableView& t, int n_keys, int n_acc, const int* ops, int match_keys, const uint64_t* src, int n_bufs,
                               uint64_t cap_rows, cudaStream_t stream) {
  if (n_bufs <= 0 || cap_rows == 0) return;
  k_table_merge_packed<<<grid_for((int64_t)cap_rows * n_bufs, kBlock, 148 * 8), kBlock, 0, stream>>>(t, n_keys, n_acc, ops, match_keys, src, n_bufs, cap_rows);
  count_launch();
  if (!match_keys && n_keys > 0) {
    k_table_merge_fixkeys<<<grid_for((int64_t)cap_rows * n_bufs, kBlock, 148 * 8), kBlock, 0, stream>>>(t, n_keys, n_acc, src, n_bufs, cap_rows);
    count_launch();
  }
  SQ_CUDA(cudaGetLastError());
}

void launch_tpch_generate(int table, int col, int64_t row_begin, int64_t n, int64_t n_customer, int flags_mode, uint64_t* dst,
                          cudaStream_t stream) {
  if (n <= 0) return;
  k_tpch_generate<<<grid_for(n, kBlock, 148 * 8), kBlock, 0, stream>>>(table, col, row_begin, n, n_customer, flags_mode, dst);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}

void launch_count_bits(const uint32_t* bitmap, int64_t n, unsigned long long* out, cudaStream_t stream) {
  if (n <= 0) return;
  k_count_bits<<<grid_for((n + 31) / 32, kBlock, 148 * 4), kBlock, 0, stream>>>(bitmap, n, out);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}

void launch_bitmap_append(uint32_t* dst, int64_t dst_bit_off, const uint32_t* src, int64_t n, cudaStream_t stream) {
  if (n <= 0) return;
  k_bitmap_append<<<grid_for((n + 63) / 32, kBlock, 148 * 4), kBlock, 0, stream>>>(dst, dst_bit_off, src, n);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}

size_t compact_num_chunks(int64_t n) { return (size_t)div_up(div_up(n, 32), kChunkWords); }

void launch_compact_count(const uint32_t* keep, int64_t n, uint32_t* chunk_counts, cudaStream_t stream) {
  if (n <= 0) return;
  int64_t chunks = (int64_t)compact_num_chunks(n);
  k_compact_count<<<grid_for(chunks, kBlock / 32, 148 * 8), kBlock, 0, stream>>>(keep, div_up(n, 32), chunks, chunk_counts);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}

void launch_scan_u32(const uint32_t* counts, int64_t m, unsigned long long* offsets, unsigned long long* total, cudaStream_t stream) {
  k_scan_u32<<<1, 1024, 0, stream>>>(counts, m, offsets, total);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}

void launch_compact_write(const uint32_t* keep, int64_t n, const unsigned long long* chunk_offsets, uint32_t* out_idx, cudaStream_t stream) {
  if (n <= 0) return;
  int64_t chunks = (int64_t)compact_num_chunks(n);
  k_compact_write<<<grid_for(chunks, kBlock / 32, 148 * 8), kBlock, 0, stream>>>(keep, div_up(n, 32), chunks, chunk_offsets, out_idx);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}

void launch_gather_u32idx(int width, const void* src, const uint32_t* src_valid, const uint32_t* idx, int64_t m, void* dst,
                          uint32_t* dst_valid, cudaStream_t stream) {
  launch_gather_t<uint32_t>(width, src, src_valid, idx, m, dst, dst_valid, stream);
}
void launch_gather_i64idx(int width, const void* src, const uint32_t* src_valid, const int64_t* idx, int64_t m, void* dst,
                          uint32_t* dst_valid, cudaStream_t stream) {
  launch_gather_t<int64_t>(width, src, src_valid, idx, m, dst, dst_valid, stream);
}

void launch_fill_u64(uint64_t* dst, int64_t n, uint64_t value, cudaStream_t stream) {
  if (n <= 0) return;
  k_fill_u64<<<grid_for(n, kBlock, 148 * 8), kBlock, 0, stream>>>(dst, n, value);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}

void launch_iota_filter_valid(const int64_t* idx, int64_t m, uint32_t* valid_out, cudaStream_t stream) {
  if (m <= 0) return;
  k_idx_valid<<<grid_for(m, kBlock), kBlock, 0, stream>>>(idx, m, valid_out);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}

void launch_table_pack(const TableView& t, int n_keys, int n_acc, uint64_t* dst, uint64_t cap_rows, cudaStream_t stream, bool mark_unchecked) {
  k_table_pack<<<grid_for(t.capacity, kBlock, 148 * 8), kBlock, 0, stream>>>(t, n_keys, n_acc, dst, cap_rows, mark_unchecked ? 1 : 0);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}

// the n groups of `t` packed in ascending first-row order (= the reference's first-appearance order):
// list occupied slots -> radix sort by min_row (CUB, a library sort of n small keys) -> ordered pack
void table_pack_sorted(const TableView& t, int n_keys, int n_acc, uint32_t n, uint64_t* dst, cudaStream_t stream, const uint32_t* slot_list,
                       int key_bits) {
  if (n == 0) return;
  uint64_t *k_in = nullptr, *k_out = nullptr;
  uint32_t *v_in = nullptr, *v_out = nullptr, *count = nullptr;
  k_in = (decltype(k_in))scratch_alloc((size_t)n * 8, stream);
  k_out = (decltype(k_out))scratch_alloc((size_t)n * 8, stream);
  v_in = (decltype(v_in))scratch_alloc((size_t)n * 4, stream);
  v_out = (decltype(v_out))scratch_alloc((size_t)n * 4, stream);
  count = (decltype(count))scratch_alloc(4, stream);
  SQ_CUDA(cudaMemsetAsync(count, 0, 4, stream));
  if (slot_list) k_table_list_from_slots<<<grid_for(n, kBlock, 148 * 8), kBlock, 0, stream>>>(t, slot_list, n, k_in, v_in);
  else k_table_list<<<grid_for(t.capacity, kBlock, 148 * 8), kBlock, 0, stream>>>(t, k_in, v_in, n, count);
  count_launch();
  if (key_bits < 1 || key_bits > 64) key_bits = 64;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, v_out, (int)n, 0, key_bits, stream);
  void* tmp = nullptr;
  tmp = (decltype(tmp))scratch_alloc(tmp_bytes ? tmp_bytes : 16, stream);
  SQ_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, (int)n, 0, key_bits, stream));
  count_launch(2 + (key_bits + 7) / 8);
  k_table_pack_ordered<<<grid_for(n, kBlock, 148 * 8), kBlock, 0, stream>>>(t, n_keys, n_acc, v_out, n, dst);
  count_launch();
  SQ_CUDA(cudaGetLastError());
  scratch_free(tmp, stream);
  scratch_free(k_in, stream);
  scratch_free(k_out, stream);
  scratch_free(v_in, stream);
  scratch_free(v_out, stream);
  scratch_free(count, stream);
}

__global__ void __launch_bounds__(kBlock) k_str_rank(const int64_t* __restrict__ ids, int64_t n, const int32_t* __restrict__ rank, int64_t* __restrict__ out,
                                                      int pack) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t id = ids[i];
    const int64_t r = (int64_t)rank[id];
    out[i] = pack ? (int64_t)(((uint64_t)r << 32) | (uint64_t)(uint32_t)id) : r;
  }
}
__global__ void __launch_bounds__(kBlock) k_str_rerank(uint64_t* __restrict__ words, const uint32_t* __restrict__ state, uint32_t capacity,
                                                        const int32_t* __restrict__ rank, uint64_t identity) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < capacity; s += stride) {
    if (state[s] != 2u) continue;
    const uint64_t v = words[s];
    if (v == identity) continue;
    const uint64_t id = v & 0xffffffffULL;
    words[s] = ((uint64_t)rank[id] << 32) | id;
  }
}
void launch_str_rank(const int64_t* ids, int64_t n, const int32_t* rank, int64_t* out, cudaStream_t stream) {
  if (n <= 0) return;
  k_str_rank<<<grid_for(n, kBlock, 148 * 8), kBlock, 0, stream>>>(ids, n, rank, out, 0);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}
void launch_str_pack(const int64_t* ids, int64_t n, const int32_t* rank, int64_t* out, cudaStream_t stream) {
  if (n <= 0) return;
  k_str_rank<<<grid_for(n, kBlock, 148 * 8), kBlock, 0, stream>>>(ids, n, rank, out, 1);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}
void launch_str_rerank(uint64_t* words, const uint32_t* state, uint32_t capacity, const int32_t* rank, uint64_t identity, cudaStream_t stream) {
  k_str_rerank<<<grid_for(capacity, kBlock, 148 * 8), kBlock, 0, stream>>>(words, state, capacity, rank, identity);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}

void launch_table_pack_partitioned(const TableView& t, int n_keys, int n_acc, uint64_t* dst, int n_parts, uint64_t cap_rows, cudaStream_t stream) {
  k_table_pack_partitioned<<<grid_for(t.capacity, kBlock, 148 * 8), kBlock, 0, stream>>>(t, n_keys, n_acc, dst, (unsigned)n_parts, cap_rows);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}
void launch_table_pack_list(const TableView& t, int n_keys, int n_acc, const uint32_t* slot_list, uint32_t n, uint64_t* dst, cudaStream_t stream) {
  if (n == 0) return;
  k_table_pack_ordered<<<grid_for(n, kBlock, 148 * 8), kBlock, 0, stream>>>(t, n_keys, n_acc, slot_list, n, dst);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}

void launch_table_rehash(const TableView& from, const TableView& to, int n_keys, int n_acc, cudaStream_t stream) {
  k_table_rehash<<<grid_for(from.capacity, kBlock, 148 * 8), kBlock, 0, stream>>>(from, to, n_keys, n_acc);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}

}  // namespace sq
