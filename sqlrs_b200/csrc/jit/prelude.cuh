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
// one 128-byte line of a column stream into L2 ahead of its use (no register is tied up, unlike an early load)
__device__ __forceinline__ void sq_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- L2 residency control.  The fused join kernels mix three kinds of traffic through the 126 MB L2: a column stream (GBs,
// no reuse: evict-first loads above), random single-use accesses to hash tables far larger than the L2 (hundreds of MB), and
// small hot structures — the Bloom filters, a few MB to tens of MB — that every probe row reads.  Measured (profiles/
// r02b_*): with the default policy the table traffic turns the whole L2 over every ~20 us and evicts the Bloom words
// between touches (90 % of the Bloom atomics and 45 % of the Bloom loads went to DRAM).  So: hot structures are accessed
// with an evict-last policy, single-use table accesses with evict-first.
__device__ __forceinline__ u64 sq_l2_evict_last() {
  u64 p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ u64 sq_l2_evict_first() {
  u64 p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ u32 sq_ld_u32_l2(const u32* a, u64 pol) {  // read-only data
  u32 v;
  asm("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ ulonglong2 sq_ld_u64x2_l2(const ulonglong2* a, u64 pol) {  // read-only data
  ulonglong2 v;
  asm("ld.global.nc.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(a), "l"(pol));
  return v;
}
// (atom.cas takes no cache hint — ptxas: "Illegal modifier '.L2::cache_hint'" — so a table insert is a plain CAS followed by a
// hinted store of the payload, which leaves the line marked evict-first)
__device__ __forceinline__ u64 sq_cas_u64_l2(u64* a, u64 cmp, u64 val, u64) { return atomicCAS(a, cmp, val); }
__device__ __forceinline__ void sq_st_u64_l2(u64* a, u64 val, u64 pol) {
  asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(a), "l"(val), "l"(pol) : "memory");
}
__device__ __forceinline__ void sq_red_or_u32_l2(u32* a, u32 val, u64 pol) {
  asm volatile("red.global.or.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(a), "r"(val), "l"(pol) : "memory");
}

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
