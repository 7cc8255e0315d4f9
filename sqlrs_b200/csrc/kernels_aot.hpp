// sqlrs_b200 — launchers of the ahead-of-time compiled kernels (kernels_aot.cu): everything that
// does not depend on a plan's expressions.  All launches are asynchronous on `stream`.
#pragma once
#include "common.hpp"

namespace sq {

// ---- synthetic TPC-H-shaped tables straight into HBM (include/sqlrs_tpch_spec.h)
void launch_tpch_generate(int table, int col, int64_t row_begin, int64_t n, int64_t n_customer, int flags_mode, uint64_t* dst,
                          cudaStream_t stream);

// ---- bitmaps
// number of set bits among the first n bits -> *out (device u64, accumulated with atomicAdd: zero it first)
void launch_count_bits(const uint32_t* bitmap, int64_t n, unsigned long long* out, cudaStream_t stream);
// dst[dst_bit_off .. +n) = src[0 .. n) (src == nullptr: all ones); dst must be zero-initialised there
void launch_bitmap_append(uint32_t* dst, int64_t dst_bit_off, const uint32_t* src, int64_t n, cudaStream_t stream);

// ---- stream compaction: keep-bitmap -> ascending list of kept row ids (warp ballot/popc)
// chunk_counts: u32[ceil(n/2048)]; chunk_offsets: u64[same]; total: u64
size_t compact_num_chunks(int64_t n);
void launch_compact_count(const uint32_t* keep, int64_t n, uint32_t* chunk_counts, cudaStream_t stream);
void launch_scan_u32(const uint32_t* counts, int64_t m, unsigned long long* offsets, unsigned long long* total, cudaStream_t stream);
void launch_compact_write(const uint32_t* keep, int64_t n, const unsigned long long* chunk_offsets, uint32_t* out_idx,
                          cudaStream_t stream);

// ---- gather (arrow compute::take): index < 0 (i64) is a NULL index -> NULL row
void launch_gather_u32idx(int width, const void* src, const uint32_t* src_valid, const uint32_t* idx, int64_t m, void* dst,
                          uint32_t* dst_valid, cudaStream_t stream);
void launch_gather_i64idx(int width, const void* src, const uint32_t* src_valid, const int64_t* idx, int64_t m, void* dst,
                          uint32_t* dst_valid, cudaStream_t stream);

// ---- Utf8 (string pool ids, device.hpp): order-dependent operators go through the pool's rank table (int32 per id)
// out[i] = rank[ids[i]]                                   (ORDER BY a Utf8 column: sort the ranks)
void launch_str_rank(const int64_t* ids, int64_t n, const int32_t* rank, int64_t* out, cudaStream_t stream);
// out[i] = rank[ids[i]] << 32 | ids[i]                    (MIN / MAX over a Utf8 column: an Int64 whose order is the strings')
void launch_str_pack(const int64_t* ids, int64_t n, const int32_t* rank, int64_t* out, cudaStream_t stream);
// words[s] = rank[id] << 32 | id with id = low half of words[s], for occupied slots whose word is not `identity`: the
// accumulated MIN / MAX strings re-expressed in the CURRENT ranks (the pool grew since the last batch)
void launch_str_rerank(uint64_t* words, const uint32_t* state, uint32_t capacity, const int32_t* rank, uint64_t identity, cudaStream_t stream);

// ---- misc
void launch_fill_u64(uint64_t* dst, int64_t n, uint64_t value, cudaStream_t stream);
void launch_iota_filter_valid(const int64_t* idx, int64_t m, uint32_t* valid_out, cudaStream_t stream);

// ---- group table maintenance (layout: struct SqTable of csrc/jit/agg.cuh)
struct TableView {
  uint32_t* state;
  uint64_t* hash;
  uint64_t* min_row;
  uint64_t* keys;
  uint32_t* knull;
  uint64_t* acc;
  uint32_t* new_slots;
  uint32_t* counters;
  uint32_t capacity;
};
// Packed row-major form of a group table: row 0 = header {group count, ...}, row 1+i =
// [hash, min_row, knull, key bits x K, accumulator words x W] (3+K+W u64 words per row).
// dst must hold (cap_rows + 1) rows and have its header zeroed; groups beyond cap_rows are counted, not written.
// mark_unchecked: see k_table_pack (a table filled by a launch whose status bits nobody has read yet)
void launch_table_pack(const TableView& t, int n_keys, int n_acc, uint64_t* dst, uint64_t cap_rows, cudaStream_t stream, bool mark_unchecked = false);
// the n groups packed in ascending first-row order (dst: (n + 1) rows; header written)
// slot_list (optional): a complete device list of the n occupied slots — skips the scan over the capacity;
// key_bits: number of significant bits of the min_row keys (fewer radix passes)
void table_pack_sorted(const TableView& t, int n_keys, int n_acc, uint32_t n, uint64_t* dst, cudaStream_t stream,
                       const uint32_t* slot_list = nullptr, int key_bits = 64);
// partial -> final merge of n_bufs packed buffers ((cap_rows + 1) rows each) into `t`; ops[w] (device):
// 0 add u64, 1 add f64, 2 min i64, 3 max i64, 4 (epoch << 40 | count): later epoch wins, equal epochs add
void launch_table_merge_packed(const TableView& t, int n_keys, int n_acc, const int* ops, int match_keys, const uint64_t* src, int n_bufs,
                               uint64_t cap_rows, cudaStream_t stream);
// re-insert every occupied slot of `from` into the (empty, initialised) table `to`
void launch_table_rehash(const TableView& from, const TableView& to, int n_keys, int n_acc, cudaStream_t stream);

}  // namespace sq

// =================================================================== hash join (kernels_join.cu)
namespace sq {

// Open-addressed table over the DISTINCT build keys; each slot owns a contiguous, ascending range
// of build row ids (CSR), so a probe emits its matches already in build insertion order.
struct JoinTableView {
  int64_t* slot_rep;      // representative build row of the slot's key, -1 = empty
  uint32_t* slot_count;   // build rows with that key
  uint64_t* slot_start;   // first position of the slot's range in `rows`
  int64_t* rows;          // build row ids grouped by slot, ascending within a slot
  uint32_t capacity;      // power of two
  const uint64_t* h;      // [n_build] row hashes of the build side (create_hashes)
  const uint64_t* keys;   // [n_keys][n_build] raw key bits (match_keys only)
  const uint32_t* knull;  // [n_build] null mask of the key tuple (match_keys only)
  int64_t n_build;
  int n_keys;
  int match_keys;         // SQLRS_MATCH_HASH_AND_KEY: compare key tuples, NULL never joins
  const uint32_t* build_keep;  // optional bitmap: build rows that pass the Filter fused below the join (nullptr = all)
  // blocked Bloom filter over the row hashes of the inserted build rows: 3 bits in ONE 32-bit word per key, >= 16 bits
  // of filter per key (join_bloom_word / join_bloom_bits below: the word comes from the hash's high half, the bit positions
  // from its low 15 bits — a dozen 32-bit instructions per probe row, which matters because the fused probe kernels are
  // instruction-issue bound).  A few MB, so it stays L2-resident under a streaming probe scan and turns a probe miss —
  // the common case of a selective join — into one 4-byte L2 read.
  uint32_t* bloom;
  uint32_t bloom_mask;    // number of words - 1 (power of two)
  int unique;             // every build key occurs once (primary-key side): slot_rep is the whole match list,
                          // slot_start / rows are not built
  // Key-in-slot layout ("kv", cuCollections-style static map) for the common single-key join compared by value
  // (n_keys == 1, match_keys): kv[2*s] = raw key bits (kJoinKvEmpty = empty slot), kv[2*s+1] = representative (smallest)
  // build row.  A probe step is ONE 16-byte read instead of the dependent chain slot_rep -> h[rep] -> keys[rep], and the
  // table needs no per-build-row hash / key arrays — which is what lets a probe kernel build the NEXT join's table
  // directly (csrc/jit/joinchain.cuh).  In this mode slot_rep = kv + 1 and rep_stride = 2.  A build key whose bits
  // equal kJoinKvEmpty cannot be stored: the insert kernels flag it and the host falls back to the slot_rep layout.
  uint64_t* kv;
  int kv_dtype;           // dtype of the key: a probe key of another type never matches (the placement hash is typed too)
  int rep_stride;         // 1, or 2 in kv mode
  int64_t n_inserted;     // host-side number (or, while a run is unvalidated, estimate) of build rows in the table
};
constexpr uint64_t kJoinKvEmpty = 0xffffffffffffffffULL;

SQ_HD inline uint32_t join_bloom_word(uint64_t h, uint32_t mask) { return (uint32_t)(h >> 32) & mask; }
SQ_HD inline uint32_t join_bloom_bits(uint64_t h) {
  const uint32_t lo = (uint32_t)h;
  return (1u << (lo & 31u)) | (1u << ((lo >> 5) & 31u)) | (1u << ((lo >> 10) & 31u));
}
uint32_t join_bloom_words(int64_t n_build);  // sizing rule: 32-bit words, power of two >= n_build / 2, within [1024, 2^25]

void launch_scan_u32_large(const uint32_t* counts, int64_t m, unsigned long long* offsets, unsigned long long* total,
                           unsigned long long* scratch /* >= ceil(m/4096)+1 */, cudaStream_t stream);
size_t scan_scratch_entries(int64_t m);

// inserts the distinct keys (slot_rep, Bloom filter, row_slot); *has_dups = 1 when some key occurred more than once
void launch_join_insert(const JoinTableView& t, int32_t* row_slot, uint32_t* has_dups, cudaStream_t stream);
// the same into the kv layout; misc[0] = 1 when some key repeats, misc[4] = 1 when a key equals kJoinKvEmpty (table unusable), misc[5] = 1 when the table ran full
void launch_join_insert_kv(const JoinTableView& t, int32_t* row_slot, uint32_t* misc, cudaStream_t stream);
// (only then) rows per slot into t.slot_count (zero-initialised) and the largest count into *max_count
void launch_join_count(const JoinTableView& t, const int32_t* row_slot, uint32_t* max_count, cudaStream_t stream);
void launch_join_fill(const JoinTableView& t, const int32_t* row_slot, uint32_t* slot_fill, cudaStream_t stream);
void launch_join_sort_ranges(const JoinTableView& t, cudaStream_t stream);
// stable fallback for heavily duplicated keys: rows = build row ids sorted by (slot, row id)
void join_fill_sorted(const JoinTableView& t, const int32_t* row_slot, cudaStream_t stream);
// fused probe path (csrc/jit/joinprobe.cuh): slot_of[r] (>= 0 slot, -1 kept but unmatched, -2 dropped by the fused Filter)
// + exclusive offsets of the 2048-row chunks -> (build row, probe row) pairs in the reference's order
constexpr int kProbeChunk = 2048;
void launch_join_probe_emit(const JoinTableView& t, const int32_t* slot_of, const unsigned long long* chunk_offsets, int64_t n_probe,
                            int keep_unmatched, int64_t* li, uint32_t* ri, cudaStream_t stream);
// bitmap of the probe rows the fused Filter kept: bit r = slot_of[r] != -2
void launch_slot_keep_bitmap(const int32_t* slot_of, int64_t n, uint32_t* bitmap, cudaStream_t stream);
// bitmap[idx[k]] = 1 for every k (idx < 0 skipped)
void launch_mark_bits_i64(const int64_t* idx, int64_t m, uint32_t* bitmap, cudaStream_t stream);
void launch_mark_bits_u32(const uint32_t* idx, int64_t m, uint32_t* bitmap, cudaStream_t stream);
// dst = ~src over the first n bits (bits past n cleared)
// (and_mask != nullptr: dst = and_mask & ~src)
void launch_bitmap_not(const uint32_t* src, int64_t n, uint32_t* dst, cudaStream_t stream, const uint32_t* and_mask = nullptr);

}  // namespace sq

// =================================================================== order / limit / finalisation (kernels_sort.cu)
namespace sq {

void launch_iota_u32(uint32_t* dst, int64_t n, uint32_t first, cudaStream_t stream);
// One STABLE pass of an LSD lexicographic sort: re-orders perm (u32[n] row ids) by one sort column — order-preserving
// 64-bit image of the value (complemented when descending), NULLs first.  Call for the LAST sort expression first.
// reverse_nulls: the run of NULL rows comes out in reverse row order (arrow's single-column descending sort).
void sort_pass(int dtype, const void* data, const uint32_t* valid, int64_t n, bool descending, bool reverse_nulls, uint32_t* perm,
               cudaStream_t stream);

// One output column of an aggregate's result, read from the packed group rows
// [hash, min_row, knull, key bits x K, accumulator words x W] (launch_table_pack / table_pack_sorted).
struct FinalizeCol {
  void* data;           // device: int64/float64 8 B, int32 4 B per row, Boolean bit-packed
  uint32_t* valid;      // device validity bitmap or nullptr
  int dtype;
  int word;             // value word within the packed row
  int null_bit;         // >= 0: group key k, NULL when bit k of the row's null mask is set
  int nvalid_word;      // >= 0: NULL when that word (number of non-NULL inputs) is 0
  int f64_sortable;     // value is the total-order integer image of a double (MIN/MAX over Float64)
  int utf8_packed;      // MIN / MAX over Utf8: the word is rank << 32 | string pool id; the column receives the id
  int count_epoch;      // COUNT under the overwrite quirk K1: (batch epoch << 40) | count
  uint64_t simple_epoch;  // SimpleAgg + K1: only the last batch counts (0 = not applicable)
};
// (the column descriptors travel as kernel parameters: no H2D copy, nothing for the host to keep alive)
void launch_finalize_groups(const uint64_t* packed, int words, int64_t n, int n_cols, const FinalizeCol* cols_host, cudaStream_t stream);
// the same into n_parts regions of (cap_rows + 1) rows each, by identity hash mod n_parts (headers must be zeroed)
void launch_table_pack_partitioned(const TableView& t, int n_keys, int n_acc, uint64_t* dst, int n_parts, uint64_t cap_rows, cudaStream_t stream);
// packed rows in the order of a given slot list (no ordering: the consumer orders by something else anyway)
void launch_table_pack_list(const TableView& t, int n_keys, int n_acc, const uint32_t* slot_list, uint32_t n, uint64_t* dst, cudaStream_t stream);

// ORDER BY ... LIMIT k without sorting: the k smallest rows under the lexicographic order of up to kTopKMaxKeys sort columns
// (same order-preserving images as sort_pass: NULLs first, descending = complemented), ties broken by `tiebreak[row]`
// (unique per row; nullptr = the row number, i.e. the stable order) -> perm_out[0 .. min(k, n)) in sorted order.
// Two launches: every CTA selects the k smallest rows of its slice by k rounds of "smallest key above the previous
// winner", one CTA does the same over the candidates.  Reads each sort column k times out of L1/L2, writes k row ids.
constexpr int kTopKMaxKeys = 4;
constexpr int kTopKMaxRows = 128;
struct TopKKeys {
  int m;
  int dtype[kTopKMaxKeys];
  int descending[kTopKMaxKeys];
  const void* data[kTopKMaxKeys];
  const uint32_t* valid[kTopKMaxKeys];
  const uint64_t* tiebreak;
};
void launch_topk(const TopKKeys& keys, int64_t n, int k, uint32_t* perm_out, cudaStream_t stream);

}  // namespace sq
