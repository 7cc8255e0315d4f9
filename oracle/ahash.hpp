// ORACLE — TEST INFRASTRUCTURE ONLY.
// Row hashing exactly as src/executor/aggregate/hash_utils.rs:13-16,161-220 with
// `RandomState::with_seeds(0, 0, 0, 0)` (hash_agg.rs:35, hash_join.rs:155).
//
// Third-party algorithm absent from /root/reference: ahash 0.8.0 (Cargo.lock:6-7),
// fallback (non-AES) hasher.  Restated from the published algorithm; the constants and
// the finishing step are PINNED by the reference's own known-answer test
// (hash_utils.rs:229-247), which tests/test_oracle_golden.py reproduces:
//   buffer = folded_multiply(value ^ k0, MULTIPLE)
//   hash   = rotate_left(folded_multiply(buffer, k1), buffer & 63)
// Pinned for 64-bit values (Int64 and Float64-as-bits).  Int32 goes through the same
// path zero-extended (Hasher::write_i32 -> write_u32 -> u64), Boolean as 0/1 — both
// UNPINNED (no reference vector).  Utf8 uses an oracle-private string hash — UNPINNED;
// hashes never leave the operators, so this only matters under 64-bit collisions.
#pragma once
#include "columns.hpp"

namespace oracle {

constexpr uint64_t AHASH_K0 = 0x452821e638d01377ULL;
constexpr uint64_t AHASH_K1 = 0xbe5466cf34e90c6cULL;
constexpr uint64_t AHASH_MULTIPLE = 6364136223846793005ULL;

inline uint64_t folded_multiply(uint64_t a, uint64_t b) {
  unsigned __int128 r = (unsigned __int128)a * (unsigned __int128)b;
  return (uint64_t)r ^ (uint64_t)(r >> 64);
}
inline uint64_t rotl64(uint64_t x, unsigned r) {
  r &= 63;
  return r ? (x << r) | (x >> (64 - r)) : x;
}
inline uint64_t hash_one_u64(uint64_t v) {
  uint64_t buf = folded_multiply(v ^ AHASH_K0, AHASH_MULTIPLE);
  return rotl64(folded_multiply(buf, AHASH_K1), (unsigned)(buf & 63));
}
// hash_utils.rs:13-16
inline uint64_t combine_hashes(uint64_t l, uint64_t r) {
  uint64_t h = (uint64_t)(17 * 37) + l;
  return h * 37 + r;
}
inline uint64_t hash_string_unpinned(const std::string& s) {
  uint64_t h = 0xcbf29ce484222325ULL;  // FNV-1a, then through the pinned finisher
  for (unsigned char ch : s) {
    h ^= ch;
    h *= 0x100000001b3ULL;
  }
  return hash_one_u64(h ^ ((uint64_t)s.size() << 56));
}

inline uint64_t hash_cell(const Column& c, int64_t r) {
  switch (c.dtype) {
    case SQLRS_DT_INT32: return hash_one_u64((uint64_t)(uint32_t)(int32_t)c.i[r]);
    case SQLRS_DT_INT64: return hash_one_u64((uint64_t)c.i[r]);
    case SQLRS_DT_BOOL: return hash_one_u64((uint64_t)(c.i[r] != 0));
    case SQLRS_DT_FLOAT64: {  // hash_utils.rs:109-153: u64::from_le_bytes(value.to_le_bytes())
      uint64_t bits;
      std::memcpy(&bits, &c.f[r], 8);
      return hash_one_u64(bits);
    }
    case SQLRS_DT_UTF8: return hash_string_unpinned(c.s[r]);
  }
  fail(SQLRS_ERR_INTERNAL, std::string("Unsupported data type in hasher: ") + dtype_name(c.dtype));
}

// create_hashes, hash_utils.rs:161-220.  `hashes` must be pre-sized and zeroed by the caller
// (hash_agg.rs:76, hash_join.rs:169,214).  A NULL cell leaves the running hash untouched
// (hash_utils.rs:91-104) — quirk K3.
inline void create_hashes(const std::vector<ColPtr>& arrays, std::vector<uint64_t>& hashes) {
  bool multi_col = arrays.size() > 1;
  for (const ColPtr& colp : arrays) {
    const Column& col = *colp;
    if ((int64_t)hashes.size() != col.n && col.dtype != SQLRS_DT_NULL)
      fail(SQLRS_ERR_INTERNAL, "create_hashes: length mismatch");
    if (col.dtype == SQLRS_DT_NULL) {  // hash_null, hash_utils.rs:18-29: hash_one(&1) with 1: i32
      uint64_t h1 = hash_one_u64(1);
      for (auto& h : hashes) h = multi_col ? combine_hashes(h1, h) : h1;
      continue;
    }
    bool no_nulls = col.valid.empty();
    for (int64_t r = 0; r < col.n; r++) {
      if (!no_nulls && !col.valid[r]) continue;
      uint64_t hv = hash_cell(col, r);
      hashes[r] = multi_col ? combine_hashes(hv, hashes[r]) : hv;
    }
  }
}

}  // namespace oracle
