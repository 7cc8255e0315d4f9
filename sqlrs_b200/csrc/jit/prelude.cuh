// sqlrs_b200 JIT prelude — compiled by NVRTC in front of every specialised kernel (no headers).
// Hand-written for sm_100a; see DESIGN.md "Kernels".
typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;

#define SQ_FULL 0xffffffffu
#define SQ_EMPTY_ROW 0xffffffffffffffffULL

// ---- streaming global loads.  Input columns are read exactly once per operator, so they are
// loaded through the read-only path with an evict-first L2 policy (no reuse to protect).
__device__ __forceinline__ i64 sq_ld_i64(const void* p, i64 r) { return __ldcs(((const i64*)p) + r); }
__device__ __forceinline__ int sq_ld_i32(const void* p, i64 r) { return __ldcs(((const int*)p) + r); }
__device__ __forceinline__ double sq_ld_f64(const void* p, i64 r) { return __ldcs(((const double*)p) + r); }
// gathers (build side of a join): random access, keep them in L1/L2
__device__ __forceinline__ i64 sq_ldg_i64(const void* p, i64 r) { return __ldg(((const i64*)p) + r); }
__device__ __forceinline__ int sq_ldg_i32(const void* p, i64 r) { return __ldg(((const int*)p) + r); }
__device__ __forceinline__ double sq_ldg_f64(const void* p, i64 r) { return __ldg(((const double*)p) + r); }
__device__ __forceinline__ bool sq_ld_bit(const void* p, i64 r) { return (__ldg(((const u32*)p) + (r >> 5)) >> (r & 31)) & 1u; }

// ---- ahash 0.8.0 fallback hasher with RandomState::with_seeds(0,0,0,0), as the reference uses it
// (src/executor/aggregate/hash_utils.rs:161-220; constants pinned by its KAT :229-247):
//   buf = folded_multiply(v ^ k0, MULTIPLE);  h = rotl(folded_multiply(buf, k1), buf & 63)
__device__ __forceinline__ u64 sq_fold_mul(u64 a, u64 b) { return (a * b) ^ __umul64hi(a, b); }
__device__ __forceinline__ u64 sq_hash_one(u64 v) {
  const u64 buf = sq_fold_mul(v ^ 0x452821e638d01377ULL, 6364136223846793005ULL);
  const u64 m = sq_fold_mul(buf, 0xbe5466cf34e90c6cULL);
  const unsigned rot = (unsigned)(buf & 63ULL);
  return (m << rot) | (m >> ((64u - rot) & 63u));
}
// combine_hashes, hash_utils.rs:13-16
__device__ __forceinline__ u64 sq_combine(u64 l, u64 r) { return (629ULL + l) * 37ULL + r; }
// slot spreader for the open-addressed tables (the row hash itself is the group identity)
__device__ __forceinline__ u32 sq_mix32(u64 h) {
  h ^= h >> 33;
  h *= 0xff51afd7ed558ccdULL;
  h ^= h >> 29;
  return (u32)h;
}

// doubles kept in a form whose signed-integer order equals the floating-point order, so that
// MIN/MAX accumulate with native 64-bit integer atomics
__device__ __forceinline__ i64 sq_f64_sortable(double d) {
  i64 b = __double_as_longlong(d);
  return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
