// sqlrs_b200 — kernels of the operators that follow the hot path (Order + Limit on the device)
// and the device-side finalisation of a group table into Arrow columns.
//   OrderExecutor   reference src/executor/order.rs:26-66 (concat -> lexsort_to_indices -> take)
//   LimitExecutor   reference src/executor/limit.rs:36-79 (batch.slice)
//   HashAgg output  reference src/executor/aggregate/hash_agg.rs:126-148 (builders -> one batch)
// The sort is an LSD sequence of STABLE radix sorts (CUB DeviceRadixSort::SortPairs) over order-preserving 64-bit
// images of the sort columns, least significant sort expression first; NULLs-first is one more stable 1-bit pass.
// All of it is small next to the scans (the input is an aggregate's output); HBM-bound on 8 B keys + 4 B row ids.
#include <cub/device/device_radix_sort.cuh>

#include "kernels_aot.hpp"

namespace sq {
namespace {

constexpr int kBlock = 256;
inline unsigned grid_for(int64_t items, int64_t cap = 148 * 8) {
  int64_t g = div_up(items, kBlock);
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (unsigned)g;
}

__device__ __forceinline__ bool bit_at(const uint32_t* bm, uint64_t i) { return (bm[i >> 5] >> (i & 31)) & 1u; }

// order-preserving u64 image of a cell: unsigned comparison of the images == the reference's comparison of the
// values (arrow sort: integers by Ord, Boolean false < true, Float64 by f64::total_cmp)
__device__ __forceinline__ uint64_t sort_image(int dtype, const void* data, uint64_t r) {
  switch (dtype) {
    case SQLRS_DT_INT64: return ((const uint64_t*)data)[r] ^ 0x8000000000000000ULL;
    case SQLRS_DT_INT32: return (uint64_t)(int64_t)((const int32_t*)data)[r] ^ 0x8000000000000000ULL;
    case SQLRS_DT_BOOL: return bit_at((const uint32_t*)data, r) ? 1ULL : 0ULL;
    case SQLRS_DT_FLOAT64: {
      const uint64_t b = ((const uint64_t*)data)[r];
      return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
    }
  }
  return 0ULL;
}

// keys[i] = image of column cell at row perm[i]; descending = complemented image.  NULL cells: 0, or — for the
// reference's single-column descending sort, which reverses the run of NULL rows (arrow sort_to_indices) —
// n-1-row so that the later stable NULL-flag pass leaves them in reverse input order.
__global__ void __launch_bounds__(kBlock) k_sort_keys(int dtype, const void* __restrict__ data, const uint32_t* __restrict__ valid,
                                                      const uint32_t* __restrict__ perm, int64_t n, int descending, int reverse_nulls,
                                                      uint64_t* __restrict__ keys) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t r = perm[i];
    uint64_t k;
    if (dtype == SQLRS_DT_NULL || (valid && !bit_at(valid, r))) {
      k = reverse_nulls ? (uint64_t)(n - 1) - r : 0ULL;
    } else {
      k = sort_image(dtype, data, r);
      if (descending) k = ~k;
    }
    keys[i] = k;
  }
}

// flags[i] = 1 for a valid cell at row perm[i], 0 for NULL (NULLs first)
__global__ void __launch_bounds__(kBlock) k_valid_flags(const uint32_t* __restrict__ valid, const uint32_t* __restrict__ perm, int64_t n,
                                                        uint32_t* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) flags[i] = bit_at(valid, perm[i]) ? 1u : 0u;
}

__global__ void __launch_bounds__(kBlock) k_iota_u32(uint32_t* __restrict__ dst, int64_t n, uint32_t first) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = first + (uint32_t)i;
}

// packed, ordered group rows -> typed Arrow columns (+ validity words by warp ballot)
constexpr int kFinalizeMaxCols = 32;
struct FinalizeCols {
  FinalizeCol c[kFinalizeMaxCols];
};
__global__ void __launch_bounds__(kBlock) k_finalize_groups(const uint64_t* __restrict__ packed, int words, int64_t n, int n_cols,
                                                            const __grid_constant__ FinalizeCols cols_p) {
  const FinalizeCol* cols = cols_p.c;
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_up = (n + 31) & ~(int64_t)31;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_up; i += stride) {
    const bool inb = i < n;
    const uint64_t* row = packed + (size_t)(1 + (inb ? i : 0)) * words;
    for (int c = 0; c < n_cols; c++) {
      const FinalizeCol d = cols[c];
      uint64_t w = row[d.word];
      bool valid = true;
      if (d.null_bit >= 0) valid = !((row[2] >> d.null_bit) & 1ULL);   // group key: bit of the tuple's null mask
      if (d.nvalid_word >= 0) valid = row[d.nvalid_word] != 0ULL;      // SUM/MIN/MAX over no non-NULL input is NULL
      if (d.count_epoch) {                                             // quirk K1 word: (batch epoch << 40) | count
        const uint64_t epoch = w >> 40;
        w &= (1ULL << 40) - 1;
        if (d.simple_epoch && epoch != d.simple_epoch) w = 0;
      }
      if (d.f64_sortable) {  // MIN/MAX(Float64) accumulate in the total-order integer image
        const int64_t s = (int64_t)w;
        w = (uint64_t)(s ^ ((s >> 63) & 0x7fffffffffffffffLL));
      }
      if (d.utf8_packed) w &= 0xffffffffULL;
      if (!valid) w = 0;
      if (inb) {
        if (d.dtype == SQLRS_DT_INT32) ((uint32_t*)d.data)[i] = (uint32_t)w;
        else if (d.dtype != SQLRS_DT_BOOL) ((uint64_t*)d.data)[i] = w;
      }
      if (d.dtype == SQLRS_DT_BOOL) {
        const uint32_t bits = __ballot_sync(0xffffffffu, inb && valid && (w & 1ULL));
        if (lane == 0) ((uint32_t*)d.data)[i >> 5] = bits;
      }
      if (d.valid) {
        const uint32_t bits = __ballot_sync(0xffffffffu, inb && valid);
        if (lane == 0) d.valid[i >> 5] = bits;
      }
    }
  }
}

// small inputs (an aggregate's few groups, the per-rank rows of a distributed top-k merge): ONE launch instead of a dozen —
// a single CTA ranks every row against all others, rank = #{j : (flag_j, key_j, j) < (flag_i, key_i, i)}, which is the
// stable NULLs-first order of the radix path below
constexpr int kSmallSort = 2048;
__global__ void __launch_bounds__(1024) k_sort_pass_small(int dtype, const void* __restrict__ data, const uint32_t* __restrict__ valid,
                                                          uint32_t* __restrict__ perm, int n, int descending, int reverse_nulls) {
  __shared__ uint64_t key_s[kSmallSort];
  __shared__ uint32_t row_s[kSmallSort];
  __shared__ uint8_t flag_s[kSmallSort];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const uint64_t r = perm[i];
    const bool is_null = dtype == SQLRS_DT_NULL || (valid && !bit_at(valid, r));
    uint64_t k;
    if (is_null) {
      k = reverse_nulls ? (uint64_t)(n - 1) - r : 0ULL;
    } else {
      k = sort_image(dtype, data, r);
      if (descending) k = ~k;
    }
    key_s[i] = k;
    row_s[i] = (uint32_t)r;
    flag_s[i] = is_null ? 0 : 1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const uint64_t k = key_s[i];
    const uint8_t f = flag_s[i];
    int rank = 0;
    for (int j = 0; j < n; j++) {
      const uint64_t kj = key_s[j];
      const uint8_t fj = flag_s[j];
      rank += (fj < f) || (fj == f && (kj < k || (kj == k && j < i)));
    }
    perm[rank] = row_s[i];
  }
}

template <typename K>
void stable_sort_pairs(K* k_in, K* k_out, uint32_t* v_in, uint32_t* v_out, int64_t n, int end_bit, cudaStream_t stream) {
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, v_out, (int)n, 0, end_bit, stream);
  void* tmp = nullptr;
  tmp = (decltype(tmp))scratch_alloc(tmp_bytes ? tmp_bytes : 16, stream);
  SQ_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, v_out, (int)n, 0, end_bit, stream));
  count_launch(end_bit > 8 ? 10 : 3);
  scratch_free(tmp, stream);
}

}  // namespace

void launch_iota_u32(uint32_t* dst, int64_t n, uint32_t first, cudaStream_t stream) {
  if (n <= 0) return;
  k_iota_u32<<<grid_for(n), kBlock, 0, stream>>>(dst, n, first);
  count_launch();
  SQ_CUDA(cudaGetLastError());
}

// perm (u32[n], initialised by the caller to the identity or to a previous pass's result) is re-ordered by ONE sort
// column, stably: ties keep their current relative order.  Call for the LAST sort expression first.
void sort_pass(int dtype, const void* data, const uint32_t* valid, int64_t n, bool descending, bool reverse_nulls, uint32_t* perm,
               cudaStream_t stream) {
  if (n <= 1) return;
  if (n >= (1LL << 31)) fail(SQLRS_ERR_UNSUPPORTED, "Order: more than 2^31 rows in one sort");
  if (n <= kSmallSort) {
    k_sort_pass_small<<<1, 1024, 0, stream>>>(dtype, data, valid, perm, (int)n, descending ? 1 : 0, reverse_nulls ? 1 : 0);
    count_launch();
    SQ_CUDA(cudaGetLastError());
    return;
  }
  uint64_t *k_in = nullptr, *k_out = nullptr;
  uint32_t *p_out = nullptr;
  k_in = (decltype(k_in))scratch_alloc((size_t)n * 8, stream);
  k_out = (decltype(k_out))scratch_alloc((size_t)n * 8, stream);
  p_out = (decltype(p_out))scratch_alloc((size_t)n * 4, stream);
  k_sort_keys<<<grid_for(n), kBlock, 0, stream>>>(dtype, data, valid, perm, n, descending ? 1 : 0, reverse_nulls ? 1 : 0, k_in);
  count_launch();
  SQ_CUDA(cudaGetLastError());
  stable_sort_pairs<uint64_t>(k_in, k_out, perm, p_out, n, 64, stream);
  if (valid || dtype == SQLRS_DT_NULL) {
    // NULLs first, whatever the direction (SortOptions::default().nulls_first); an all-NULL column only needs the
    // value pass above (every key equal, or the reversed row ids)
    if (valid) {
      uint32_t* f_in = (uint32_t*)k_in;  // reuse: n * 4 <= n * 8
      uint32_t* f_out = (uint32_t*)k_out;
      k_valid_flags<<<grid_for(n), kBlock, 0, stream>>>(valid, p_out, n, f_in);
      count_launch();
      SQ_CUDA(cudaGetLastError());
      stable_sort_pairs<uint32_t>(f_in, f_out, p_out, perm, n, 1, stream);
    } else {
      SQ_CUDA(cudaMemcpyAsync(perm, p_out, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream));
    }
  } else {
    SQ_CUDA(cudaMemcpyAsync(perm, p_out, (size_t)n * 4, cudaMemcpyDeviceToDevice, stream));
  }
  scratch_free(k_in, stream);
  scratch_free(k_out, stream);
  scratch_free(p_out, stream);
}

void launch_finalize_groups(const uint64_t* packed, int words, int64_t n, int n_cols, const FinalizeCol* cols_host, cudaStream_t stream) {
  if (n <= 0 || n_cols <= 0) return;
  for (int first = 0; first < n_cols; first += kFinalizeMaxCols) {
    FinalizeCols p{};
    const int m = std::min(kFinalizeMaxCols, n_cols - first);
    for (int c = 0; c < m; c++) p.c[c] = cols_host[first + c];
    k_finalize_groups<<<grid_for(n), kBlock, 0, stream>>>(packed, words, n, m, p);
    count_launch();
    SQ_CUDA(cudaGetLastError());
  }
}

// ------------------------------------------------------------------ top-k (ORDER BY ... LIMIT k)
namespace {
// Composite key of a row under M sort columns: w[0] = tie-break word (compared LAST), w[1 + 2j] / w[2 + 2j] = (valid flag,
// order-preserving image) of column j.  The kernels are instantiated per M so the keys are exactly 2M + 1 registers wide.
template <int M>
struct TopKKey {
  uint64_t w[2 * M + 1];
};
template <int M>
__device__ __forceinline__ bool topk_less(const TopKKey<M>& a, const TopKKey<M>& b) {
  bool less = false, decided = false;
#pragma unroll
  for (int j = 1; j < 2 * M + 1; j++) {
    const bool ne = a.w[j] != b.w[j];
    if (!decided && ne) less = a.w[j] < b.w[j];
    decided = decided || ne;
  }
  if (!decided) less = a.w[0] < b.w[0];
  return less;
}
template <int M>
__device__ __forceinline__ void topk_load(const TopKKeys& keys, uint64_t row, TopKKey<M>& out) {
#pragma unroll
  for (int j = 0; j < M; j++) {
    const bool is_null = keys.valid[j] && !bit_at(keys.valid[j], row);
    uint64_t img = 0;
    if (!is_null) {
      img = sort_image(keys.dtype[j], keys.data[j], row);
      if (keys.descending[j]) img = ~img;
    }
    out.w[1 + 2 * j] = is_null ? 0ULL : 1ULL;  // NULLs first, whatever the direction
    out.w[2 + 2 * j] = img;
  }
  out.w[0] = keys.tiebreak ? keys.tiebreak[row] : row;
}

// One reduction pass, one WARP per slice of kTopKWarpRows entries (rows_in[...], or the row numbers themselves when rows_in ==
// nullptr): every lane loads the keys of its kTopKPer entries into registers ONCE, then the warp emits its k smallest in
// order — k rounds of a shuffle argmin over registers; no shared memory, no barrier, no memory traffic in the rounds.
// Passes repeat until one warp is left (launch_topk): every pass shrinks the candidates kTopKWarpRows / k-fold.
constexpr int kTopKPer = 8, kTopKWarpRows = 32 * kTopKPer;
template <int M>
__global__ void __launch_bounds__(kBlock) k_topk_pass(const __grid_constant__ TopKKeys keys, const uint32_t* __restrict__ rows_in, int64_t n_in, int k,
                                                      uint32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  const int64_t base = warp * kTopKWarpRows;
  if (base >= n_in) return;
  TopKKey<M> ck[kTopKPer];
  uint32_t crow[kTopKPer];
  bool live[kTopKPer];
#pragma unroll
  for (int c = 0; c < kTopKPer; c++) {
    const int64_t i = base + c * 32 + lane;
    crow[c] = i < n_in ? (rows_in ? rows_in[i] : (uint32_t)i) : 0xffffffffu;
    live[c] = crow[c] != 0xffffffffu;
  }
#pragma unroll
  for (int c = 0; c < kTopKPer; c++) topk_load<M>(keys, live[c] ? crow[c] : 0, ck[c]);
  uint32_t* dst = out + warp * k;
  for (int round = 0; round < k; round++) {
    TopKKey<M> best = ck[0];
    uint32_t best_row = crow[0];
    bool has = live[0];
#pragma unroll
    for (int c = 1; c < kTopKPer; c++)
      if (live[c] && (!has || topk_less<M>(ck[c], best))) {
        best = ck[c];
        best_row = crow[c];
        has = true;
      }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      TopKKey<M> o;
#pragma unroll
      for (int j = 0; j < 2 * M + 1; j++) o.w[j] = __shfl_xor_sync(0xffffffffu, best.w[j], d);
      const uint32_t orow = __shfl_xor_sync(0xffffffffu, best_row, d);
      const bool ohas = __shfl_xor_sync(0xffffffffu, has ? 1 : 0, d) != 0;
      if (ohas && (!has || topk_less<M>(o, best))) {
        best = o;
        best_row = orow;
        has = true;
      }
    }
    if (lane == 0) dst[round] = has ? best_row : 0xffffffffu;
    if (!has) {
      if (lane == 0)
        for (int j = round + 1; j < k; j++) dst[j] = 0xffffffffu;
      break;
    }
#pragma unroll
    for (int c = 0; c < kTopKPer; c++)
      if (live[c] && crow[c] == best_row) live[c] = false;
  }
}
}  // namespace

void launch_topk(const TopKKeys& keys, int64_t n, int k, uint32_t* perm_out, cudaStream_t stream) {
  if (n <= 0 || k <= 0) return;
  if (keys.m < 1 || keys.m > kTopKMaxKeys || k > kTopKMaxRows || n >= (1LL << 32) - 2) fail(SQLRS_ERR_INTERNAL, "launch_topk: unsupported shape");
  const uint32_t* in = nullptr;
  int64_t n_in = n;
  uint32_t* bufs[2] = {nullptr, nullptr};
  int which = 0;
  for (;;) {
    const int64_t warps = div_up(n_in, kTopKWarpRows);
    uint32_t* dst = perm_out;
    if (warps > 1) {
      if (!bufs[which]) bufs[which] = (uint32_t*)scratch_alloc((size_t)warps * k * 4, stream);  // (later passes need less)
      dst = bufs[which];
    }
    const unsigned grid = (unsigned)div_up(warps, kBlock / 32);
    switch (keys.m) {
      case 1: k_topk_pass<1><<<grid, kBlock, 0, stream>>>(keys, in, n_in, k, dst); break;
      case 2: k_topk_pass<2><<<grid, kBlock, 0, stream>>>(keys, in, n_in, k, dst); break;
      case 3: k_topk_pass<3><<<grid, kBlock, 0, stream>>>(keys, in, n_in, k, dst); break;
      default: k_topk_pass<4><<<grid, kBlock, 0, stream>>>(keys, in, n_in, k, dst); break;
    }
    count_launch();
    SQ_CUDA(cudaGetLastError());
    if (warps == 1) break;
    in = dst;
    n_in = warps * k;
    which ^= 1;
  }
  for (uint32_t* b : bufs)
    if (b) scratch_free(b, stream);
}

}  // namespace sq
