// sqlrs_b200 — HashJoinExecutor on the GPU (see join.hpp; kernels: kernels_join.cu, csrc/jit/eval.cuh).
#include "join.hpp"

#include <algorithm>
#include <cstring>

#include "kernels_aot.hpp"

namespace sq {

// highest load factor a key-in-slot table is sized for, in percent (capacity = the power of two that keeps the load below it;
// tuning experiments: SQLRS_B200_KV_MAXLOAD)
static uint64_t kv_max_load_pct() {
  static const uint64_t pct = [] {
    const char* e = std::getenv("SQLRS_B200_KV_MAXLOAD");
    return (uint64_t)(e ? std::min(95, std::max(10, atoi(e))) : 50);
  }();
  return pct;
}

ProbeProgram gen_probe_program(const std::vector<ColInfo>& cols, const std::vector<ExprCopy>& right_keys, const ExprCopy& probe_pred, bool jmatch) {
  std::ostringstream s;
  RowProgram p1(cols);
  std::string p1_pass = "true";
  if (!probe_pred.empty()) {
    Val pp = p1.compile(probe_pred, 0);
    if (pp.dtype != SQLRS_DT_BOOL) fail(SQLRS_ERR_INTERNAL, "filter executor expected evaluate boolean array");
    p1_pass = "(n" + std::to_string(pp.id) + " && v" + std::to_string(pp.id) + ")";
  }
  // the reference evaluates the keys on the rows the Filter kept: a failing key expression only counts there
  std::vector<Val> jkeys;
  for (const ExprCopy& e : right_keys) jkeys.push_back(p1.compile(e, probe_pred.empty() ? 0 : 1));
  std::vector<int> jraw;
  for (const Val& k : jkeys) jraw.push_back(p1.emit_raw_bits(k));
  const int jh = jmatch ? p1.emit_mix_hash(jraw, jkeys) : p1.emit_row_hash(jkeys);  // as the build side (key_request)
  const int JK = (int)jkeys.size();
  s << "#define SQ_JKEYS " << JK << "\n#define SQ_JMATCH " << (jmatch ? 1 : 0) << "\n#define SQ_JKEY0_DTYPE " << (JK > 0 ? jkeys[0].dtype : 0) << "\n";
  s << "struct SqProbe { bool pass; u64 h; u64 kb[" << std::max(JK, 1) << "]; u32 knull; };\n";
  s << "__device__ __forceinline__ void sq_probe_row(const SqIn& in, i64 r, SqProbe& p, bool& e0, bool& e1) {\n" << p1.body_str();
  s << "  p.pass = " << p1_pass << ";\n  p.h = v" << jh << ";\n";
  std::string jknull = "0u";
  for (int k = 0; k < JK; k++) {
    s << "  p.kb[" << k << "] = v" << jraw[k] << ";\n";
    jknull += " | (n" + std::to_string(jkeys[k].id) + " ? 0u : " + std::to_string(1u << k) + "u)";
  }
  s << "  p.knull = " << jknull << ";\n}\n";
  // What phase B needs of a candidate row besides its number (csrc/jit/joinagg.cuh): ONE queued u64 — the key bits of a
  // single compared key (the hash is re-derived from them), or the row hash itself when the hash is the identity.
  // Several compared keys (SQ_PQMODE 0): phase B re-evaluates the row.
  const int pqmode = jmatch ? (JK == 1 ? 1 : 0) : 2;
  s << "#define SQ_PQMODE " << pqmode << "\n";
  if (pqmode == 1) {
    s << "__device__ __forceinline__ u64 sq_probe_rehash(const u64* kb) {\n" << RowProgram::mix_hash_of_bits_source({jkeys[0].dtype}) << "}\n";
    s << "__device__ __forceinline__ u64 sq_probe_qv(const SqProbe& p) { return p.kb[0]; }\n";
    s << "__device__ __forceinline__ void sq_probe_unq(u64 qv, SqProbe& p) { p.pass = true; p.kb[0] = qv; p.knull = 0u; p.h = sq_probe_rehash(p.kb); }\n";
  } else if (pqmode == 2) {
    s << "__device__ __forceinline__ u64 sq_probe_qv(const SqProbe& p) { return p.h; }\n";
    s << "__device__ __forceinline__ void sq_probe_unq(u64 qv, SqProbe& p) { p.pass = true; p.h = qv; p.knull = 0u; }\n";
  }
  ProbeProgram out;
  // TMA variant: the same statements with the streaming column loads redirected to a shared-memory tile.  Only when
  // every loaded column is 8 bytes wide (bulk copies of whole 2048-row column slices) and at most 3 are read.
  const std::string body = p1.body_str();
  std::vector<int> tile_cols;
  bool ok = body.find("SQ_LD_I32(") == std::string::npos && body.find("SQ_LD_BOOL(") == std::string::npos;
  for (const char* macro : {"SQ_LD_I64(", "SQ_LD_F64("}) {
    for (size_t pos = body.find(macro); ok && pos != std::string::npos; pos = body.find(macro, pos + 1)) {
      const int c = atoi(body.c_str() + pos + strlen(macro));
      if (std::find(tile_cols.begin(), tile_cols.end(), c) == tile_cols.end()) tile_cols.push_back(c);
    }
  }
  // L2 prefetch of the 8-byte columns the probe program streams: `rows` rows starting at row0 (lane L takes the L-th 128-byte line)
  s << "__device__ __forceinline__ void sq_probe_prefetch(const SqIn& in, i64 row0, i64 n, int rows, int lane) {\n";
  s << "  const i64 r = row0 + (i64)lane * 16;\n  if (lane * 16 >= rows || r >= n) return;\n";
  if (ok)
    for (int c : tile_cols) s << "  sq_prefetch_l2((const char*)in.col[" << c << "] + r * 8);\n";
  s << "}\n";
  if (ok && !tile_cols.empty() && tile_cols.size() <= 3) {
    std::string tb = body;
    auto replace_all = [&](const std::string& from, const std::string& to) {
      for (size_t pos = tb.find(from); pos != std::string::npos; pos = tb.find(from, pos + to.size())) tb.replace(pos, from.size(), to);
    };
    for (size_t k = 0; k < tile_cols.size(); k++) {
      const std::string c = std::to_string(tile_cols[k]), at = "tile[" + std::to_string(k) + " * SQ_TROWS + t]";
      replace_all("SQ_LD_I64(" + c + ", r)", "((long long)" + at + ")");
      replace_all("SQ_LD_F64(" + c + ", r)", "__longlong_as_double((long long)" + at + ")");
    }
    s << "#define SQ_TMA 1\n#define SQ_TROWS " << tma_shape().tile_rows << "\n#define SQ_TSTAGES " << tma_shape().stages << "\n#define SQ_TCONSUMERS "
      << tma_shape().consumers << "\n#define SQ_TILE_NCOLS " << tile_cols.size() << "\n#define SQ_TILE_COLS {";
    for (size_t k = 0; k < tile_cols.size(); k++) s << (k ? ", " : "") << tile_cols[k];
    s << "}\n";
    s << "__device__ __forceinline__ void sq_probe_row_tile(const SqIn& in, const u64* __restrict__ tile, int t, i64 r, SqProbe& p, bool& e0, bool& e1) {\n" << tb;
    s << "  p.pass = " << p1_pass << ";\n  p.h = v" << jh << ";\n";
    for (int k = 0; k < JK; k++) s << "  p.kb[" << k << "] = v" << jraw[k] << ";\n";
    s << "  p.knull = " << jknull << ";\n}\n";
    out.tile_cols = tile_cols;
  } else {
    s << "#define SQ_TMA 0\n";
  }
  out.src = s.str();
  return out;
}

const TmaShape& tma_shape() {
  static const TmaShape shape = [] {
    TmaShape s{1024, 4, 8};
    if (const char* e = std::getenv("SQLRS_B200_TMA_TROWS")) s.tile_rows = std::max(256, atoi(e) / 256 * 256);
    if (const char* e = std::getenv("SQLRS_B200_TMA_STAGES")) s.stages = std::min(16, std::max(2, atoi(e)));
    if (const char* e = std::getenv("SQLRS_B200_TMA_CONSUMERS")) s.consumers = std::min(31, std::max(1, atoi(e)));
    while (s.tile_rows % (32 * s.consumers) != 0 && s.consumers > 1) s.consumers--;  // every consumer warp takes whole 32-row groups
    return s;
  }();
  return shape;
}

std::string gen_build_decls(const std::vector<ColInfo>& build_cols) {
  std::ostringstream s;
  const size_t nb = build_cols.size() ? build_cols.size() : 1;
  s << "struct SqInB { const void* col[" << nb << "]; const u32* val[" << nb << "]; };\n";
  s << "#define SQ_LDB_I64(c, b) sq_ldg_i64(inb.col[c], b)\n#define SQ_LDB_I32(c, b) sq_ldg_i32(inb.col[c], b)\n";
  s << "#define SQ_LDB_F64(c, b) sq_ldg_f64(inb.col[c], b)\n#define SQ_LDB_BOOL(c, b) sq_ld_bit(inb.col[c], b)\n";
  s << "#define SQ_VALIDB(c, b) sq_ld_bit(inb.val[c], b)\n";
  return s.str();
}

struct JoinOp::Impl {
  // build side
  std::vector<DBatch> left_batches;
  std::vector<std::vector<DCol>> left_key_parts;  // per batch: [hash, (raw bits x K, nullmask)]
  int64_t left_rows = 0;
  bool sealed = false;
  DBatch left_single;
  DCol h_all, knull_all;
  BufPtr keys_all;  // [K][n_build]
  // table
  BufPtr slot_rep, slot_count, slot_start, rows, bloom;
  uint32_t capacity = 0;
  BufPtr visited_left;  // bitmap over build rows (Left/Full)
  std::unique_ptr<EvalProgram> left_prog, filter_prog;
  std::map<std::string, JitKernel*> probe_kernels;  // fused probe kernels by probe-batch schema signature
  uint32_t max_count = 0;                           // largest number of build rows sharing one key
  JoinTableView view{};
  // plan-level fusion (SQLRS plan executor): Filters directly below the join run inside the key evaluation,
  // and output columns nobody above reads are not gathered
  ExprCopy build_pred, probe_pred;
  std::vector<DCol> left_keep_parts;
  DCol keep_all;
  int key0_dtype = 0;  // static dtype of the first key expression (kv layout)
  std::vector<bool> needed;  // per output field; empty = all
  std::vector<BufPtr> adopted;  // adopt_build: the buffers the view points into
  bool fused_build = false;      // enable_fused_build
  std::vector<DBatch> stashed;   // build batches not evaluated yet (fused build)
  JitKernel* build_kernel = nullptr;
};

// outputs: [hash, (raw bits x K, null mask)?, (keep mask of the fused Filter)?]
static EvalRequest key_request(const std::vector<ExprCopy>& keys, bool match_keys, const ExprCopy& pred) {
  EvalRequest r;
  for (const ExprCopy& k : keys) {
    r.exprs.push_back(k);
    r.is_key.push_back(true);
  }
  // hash-only identity needs the reference's row hash; with key comparison any placement hash does
  r.outs.push_back({match_keys ? OUT_MIXHASH : OUT_HASH, 0});
  if (match_keys) {
    for (size_t k = 0; k < keys.size(); k++) r.outs.push_back({OUT_RAWBITS, (int)k});
    r.outs.push_back({OUT_NULLMASK, 0});
  }
  if (!pred.empty()) {
    r.exprs.push_back(pred);
    r.is_key.push_back(false);
    r.outs.push_back({OUT_KEEP, (int)r.exprs.size() - 1});
  }
  return r;
}

JoinOp::JoinOp(int join_type, std::vector<ExprCopy> left_keys, std::vector<ExprCopy> right_keys, ExprCopy filter,
               std::vector<Field> out_fields, const Options& opt)
    : ctx_(opt), opt_(opt), join_type_(join_type), left_keys_(std::move(left_keys)), right_keys_(std::move(right_keys)),
      filter_(std::move(filter)), out_fields_(std::move(out_fields)), impl_(new Impl()) {
  if (left_keys_.size() != right_keys_.size() || left_keys_.empty()) fail(SQLRS_ERR_INTERNAL, "HashJoin must has on condition");
  if (left_keys_.size() > 16) fail(SQLRS_ERR_UNSUPPORTED, "more than 16 join keys");
  if (!filter_.empty()) {
    EvalRequest r;
    r.exprs.push_back(filter_);
    r.is_key.push_back(false);
    r.outs.push_back({OUT_KEEP, 0});
    impl_->filter_prog = std::make_unique<EvalProgram>(std::move(r));
  }
}
JoinOp::~JoinOp() = default;

void JoinOp::set_side_predicates(ExprCopy build_pred, ExprCopy probe_pred) {
  impl_->build_pred = std::move(build_pred);
  impl_->probe_pred = std::move(probe_pred);
}
void JoinOp::set_needed_columns(std::vector<bool> needed) { impl_->needed = std::move(needed); }

// hash_join.rs:161-181
void JoinOp::enable_fused_build() {
  const bool mk = opt_.match_mode == SQLRS_MATCH_HASH_AND_KEY;
  impl_->fused_build = mk && left_keys_.size() == 1 && (join_type_ == SQLRS_JOIN_INNER || join_type_ == SQLRS_JOIN_RIGHT) && !std::getenv("SQLRS_B200_NO_KV") &&
                       !std::getenv("SQLRS_B200_NO_FUSED_BUILD");
}

// scan + Filter + key hash + insert of the ONE stashed build batch in one kernel; false = take the materialising path
bool JoinOp::seal_fused() {
  Impl& im = *impl_;
  if (im.stashed.size() != 1 || im.stashed[0].n <= 0) return false;
  const DBatch& b = im.stashed[0];
  const int64_t n = b.n;
  std::vector<ColInfo> cols = col_infos(b);
  int key_dtype = 0;
  {
    RowProgram p(cols);
    key_dtype = p.compile(left_keys_[0], im.build_pred.empty() ? 0 : 1).dtype;
  }
  if (key_dtype == SQLRS_DT_NULL) return false;
  if (!im.build_kernel) {
    const std::string src = gen_input_decls(cols) + gen_probe_program(cols, left_keys_, im.build_pred, true).src;
    im.build_kernel = jit_get("join_table+joinbuild", src, "sq_joinbuild_kernel");
  }
  std::vector<const void*> in_blob(2 * std::max<size_t>(b.cols.size(), 1), nullptr);
  for (size_t c = 0; c < b.cols.size(); c++) {
    in_blob[c] = b.cols[c].data;
    in_blob[std::max<size_t>(b.cols.size(), 1) + c] = b.cols[c].valid;
  }
  struct BuildOut {
    uint64_t* kv;
    uint32_t* bloom;
    uint32_t capacity, bloom_mask;
    uint32_t* flags;
    unsigned long long* kept;
  };
  BufPtr status = dev_alloc_zero(ctx_, 40);  // u32 flags[8] + u64 kept
  const int sms = device_sm_count(ctx_.device);
  const int per_sm = std::max(1, jit_max_blocks_per_sm(im.build_kernel, 256, 0));
  auto launch = [&](BuildOut out) {
    int64_t n_arg = n;
    void* args[] = {in_blob.data(), &n_arg, &out};
    const unsigned grid = (unsigned)std::min<int64_t>(std::max<int64_t>(div_up(n, 256 * 8), 1), (int64_t)sms * per_sm);  // 256 threads x 8 rows per trip
    KernelEvent ev(opt_.flags, ctx_.stream, out.kv ? "sq_joinbuild_kernel" : "sq_joinbuild_kernel (count)");
    jit_launch(im.build_kernel, grid, 256, 0, ctx_.stream, args);
  };
  const bool deferred = hint_pinned_ && hint_kept_ >= 0;
  int64_t n_insert = 0;
  if (deferred) {
    n_insert = hint_kept_ + hint_kept_ / 8 + 64;
  } else {  // how many rows does the Filter keep?  One count-only pass over the Filter / key columns
    launch(BuildOut{nullptr, nullptr, 0, 0, (uint32_t*)status->p, (unsigned long long*)((uint32_t*)status->p + 8)});
    unsigned long long kept = 0;
    SQ_CUDA(cudaMemcpyAsync(&kept, (uint32_t*)status->p + 8, 8, cudaMemcpyDeviceToHost, ctx_.stream));
    SQ_CUDA(cudaStreamSynchronize(ctx_.stream));
    n_insert = (int64_t)kept;
    SQ_CUDA(cudaMemsetAsync(status->p, 0, 40, ctx_.stream));
  }
  uint64_t cap = 1024;
  while (cap * kv_max_load_pct() < 100ULL * (uint64_t)n_insert) cap <<= 1;
  if (cap > (1ULL << 30)) return false;
  BufPtr kv = dev_alloc(ctx_, cap * 16);
  SQ_CUDA(cudaMemsetAsync(kv->p, 0xff, cap * 16, ctx_.stream));
  const uint32_t bloom_words = join_bloom_words(n_insert);
  BufPtr bloom = dev_alloc_zero(ctx_, (size_t)bloom_words * 4);
  launch(BuildOut{(uint64_t*)kv->p, (uint32_t*)bloom->p, (uint32_t)cap, bloom_words - 1, (uint32_t*)status->p, (unsigned long long*)((uint32_t*)status->p + 8)});
  if (deferred) {  // unique keys assumed; the caller checks {flags[0..5], kept} once the stream has been synchronised
    SQ_CUDA(cudaMemcpyAsync(hint_pinned_, status->p, 40, cudaMemcpyDeviceToHost, ctx_.stream));
    sealed_deferred_ = true;
  } else {
    uint32_t flags[10] = {0};
    SQ_CUDA(cudaMemcpyAsync(flags, status->p, 40, cudaMemcpyDeviceToHost, ctx_.stream));
    SQ_CUDA(cudaStreamSynchronize(ctx_.stream));
    if (flags[3]) fail(SQLRS_ERR_ARROW, std::string(arithmetic_error_text()) + " (join key)");
    if (flags[0] || flags[4] || flags[5]) return false;  // repeated key / unrepresentable key: the materialising path handles those
    if (hint_pinned_) std::memcpy(hint_pinned_, flags, 40);
  }
  ctx_.defer([status]() {});
  im.left_single = b;
  im.slot_rep = kv;
  im.bloom = bloom;
  im.capacity = (uint32_t)cap;
  im.max_count = 1;
  JoinTableView& v = im.view;
  v = JoinTableView{};
  v.capacity = (uint32_t)cap;
  v.kv = (uint64_t*)kv->p;
  v.kv_dtype = key_dtype;
  v.slot_rep = (int64_t*)kv->p + 1;
  v.rep_stride = 2;
  v.n_build = n;
  v.n_keys = 1;
  v.match_keys = 1;
  v.bloom = (uint32_t*)bloom->p;
  v.bloom_mask = bloom_words - 1;
  v.unique = 1;
  v.n_inserted = n_insert;
  im.stashed.clear();
  return true;
}

void JoinOp::build_push(const DBatch& batch) {
  Trace tr("join.build_push", ctx_.stream);
  if (impl_->sealed) fail(SQLRS_ERR_INVALID_ARG, "hash_join: build_push after probe");
  ctx_.reap();
  if (impl_->fused_build) {
    impl_->stashed.push_back(batch);
    impl_->left_rows += batch.n;
    return;
  }
  eval_push(batch);
}

void JoinOp::eval_push(const DBatch& batch) {
  const bool mk = opt_.match_mode == SQLRS_MATCH_HASH_AND_KEY;
  if (!impl_->left_prog) impl_->left_prog = std::make_unique<EvalProgram>(key_request(left_keys_, mk, impl_->build_pred));
  EvalResult keys;
  {
    KernelEvent ev(opt_.flags, ctx_.stream, "sq_eval_kernel (join build keys + fused Filter)");
    keys = impl_->left_prog->run(ctx_, batch, "join key");
  }
  if (!keys.expr_dtypes.empty()) impl_->key0_dtype = keys.expr_dtypes[0];
  if (!impl_->build_pred.empty()) {
    impl_->left_keep_parts.push_back(keys.cols.back());
    keys.cols.pop_back();
  }
  impl_->left_key_parts.push_back(keys.cols);
  impl_->left_batches.push_back(batch);
  impl_->left_rows += batch.n;
}

// concat_batches (:187) + the hash table over the build side
void JoinOp::seal() {
  Impl& im = *impl_;
  if (im.sealed) return;
  Trace tr("join.seal", ctx_.stream);
  im.sealed = true;
  if (!im.stashed.empty()) {
    if (seal_fused()) return;
    std::vector<DBatch> batches = std::move(im.stashed);  // the materialising path after all
    im.stashed.clear();
    im.left_rows = 0;
    sealed_deferred_ = false;
    for (const DBatch& b : batches) eval_push(b);
  }
  if (im.left_batches.empty()) return;
  const int64_t n = im.left_rows;
  const int K = (int)left_keys_.size();
  const bool mk = opt_.match_mode == SQLRS_MATCH_HASH_AND_KEY;
  im.left_single.fields = im.left_batches[0].fields;
  im.left_single.n = n;
  for (size_t c = 0; c < im.left_single.fields.size(); c++) {
    std::vector<DCol> parts;
    for (const DBatch& b : im.left_batches) {
      if (b.cols.size() != im.left_single.fields.size()) fail(SQLRS_ERR_ARROW, "concat_batches: schema mismatch");
      parts.push_back(b.cols[c]);
    }
    im.left_single.cols.push_back(concat_cols(ctx_, parts, im.left_batches[0].cols[c].dtype));
  }
  {
    std::vector<DCol> parts;
    for (auto& p : im.left_key_parts) parts.push_back(p[0]);
    im.h_all = concat_cols(ctx_, parts, SQLRS_DT_INT64);
  }
  if (mk) {
    im.keys_all = dev_alloc(ctx_, (size_t)std::max<int64_t>(n, 1) * 8 * K);
    for (int k = 0; k < K; k++) {
      int64_t off = 0;
      for (auto& p : im.left_key_parts) {
        if (p[1 + k].n)
          SQ_CUDA(cudaMemcpyAsync((uint64_t*)im.keys_all->p + (size_t)k * n + off, p[1 + k].data, (size_t)p[1 + k].n * 8,
                                  cudaMemcpyDeviceToDevice, ctx_.stream));
        off += p[1 + k].n;
      }
    }
    std::vector<DCol> parts;
    for (auto& p : im.left_key_parts) parts.push_back(p[1 + K]);
    im.knull_all = concat_cols(ctx_, parts, SQLRS_DT_INT32);
  }
  if (!im.build_pred.empty()) im.keep_all = concat_cols(ctx_, im.left_keep_parts, SQLRS_DT_BOOL);
  im.left_batches.clear();
  im.left_key_parts.clear();
  im.left_keep_parts.clear();

  // table: capacity >= 2 x the build rows that can be inserted.  With a Filter fused below the build side that is the
  // number of kept rows (one bit count + sync): Q3's customer table shrinks 5x, which is what keeps it L2-resident
  // under the probe scan
  int64_t n_insert = n;
  bool use_kv = mk && K == 1 && !std::getenv("SQLRS_B200_NO_KV");
  const bool deferred = use_kv && hint_pinned_ && n > 0 && (im.build_pred.empty() || hint_kept_ >= 0);
  if (!im.build_pred.empty() && n > 0) {
    BufPtr cnt = dev_alloc_zero(ctx_, 8);
    launch_count_bits((const uint32_t*)im.keep_all.data, n, (unsigned long long*)cnt->p, ctx_.stream);
    if (deferred) {  // sized from the previous run's count (+12 %); this run's count travels to the caller for the next one
      SQ_CUDA(cudaMemcpyAsync(hint_pinned_ + 8, cnt->p, 8, cudaMemcpyDeviceToHost, ctx_.stream));
      n_insert = hint_kept_ + hint_kept_ / 8 + 64;
      ctx_.defer([cnt]() {});
    } else {
      unsigned long long kept = 0;
      SQ_CUDA(cudaMemcpyAsync(&kept, cnt->p, 8, cudaMemcpyDeviceToHost, ctx_.stream));
      SQ_CUDA(cudaStreamSynchronize(ctx_.stream));
      n_insert = (int64_t)kept;
      if (hint_pinned_) *(unsigned long long*)(hint_pinned_ + 8) = kept;
    }
  } else if (hint_pinned_) {
    *(unsigned long long*)(hint_pinned_ + 8) = (unsigned long long)n;
  }
  uint64_t cap = 1024;
  while (cap * kv_max_load_pct() < 100ULL * (uint64_t)n_insert) cap <<= 1;
  if (cap > (1ULL << 30)) fail(SQLRS_ERR_UNSUPPORTED, "join build side too large for one table (> 2^29 rows)");
  im.capacity = (uint32_t)cap;
  JoinTableView& v = im.view;
  v = JoinTableView{};
  v.slot_count = nullptr;  // per-slot counts and the CSR row lists are only built when some key repeats (below)
  v.slot_start = nullptr;
  v.rows = nullptr;
  v.unique = 1;
  v.capacity = im.capacity;
  v.h = (const uint64_t*)im.h_all.data;
  v.keys = mk ? (const uint64_t*)im.keys_all->p : nullptr;
  v.knull = mk ? (const uint32_t*)im.knull_all.data : nullptr;
  v.n_build = n;
  v.n_keys = K;
  v.match_keys = mk ? 1 : 0;
  v.build_keep = im.build_pred.empty() ? nullptr : (const uint32_t*)im.keep_all.data;
  v.n_inserted = n_insert;
  const uint32_t bloom_words = join_bloom_words(n_insert);
  im.bloom = dev_alloc_zero(ctx_, (size_t)bloom_words * 4);
  v.bloom = (uint32_t*)im.bloom->p;
  v.bloom_mask = bloom_words - 1;
  im.max_count = n_insert > 0 ? 1 : 0;
  // single key compared by value: key-in-slot layout (kernels_aot.hpp).  A key whose bits equal the empty marker cannot
  // be stored there; the insert kernel flags it and the table is rebuilt in the slot_rep layout.
  for (int attempt = 0; attempt < 2; attempt++) {
    if (use_kv) {
      im.slot_rep = dev_alloc(ctx_, cap * 16);
      SQ_CUDA(cudaMemsetAsync(im.slot_rep->p, 0xff, cap * 16, ctx_.stream));
      v.kv = (uint64_t*)im.slot_rep->p;
      v.kv_dtype = im.key0_dtype;
      v.slot_rep = (int64_t*)im.slot_rep->p + 1;
      v.rep_stride = 2;
    } else {
      im.slot_rep = dev_alloc(ctx_, cap * 8);
      SQ_CUDA(cudaMemsetAsync(im.slot_rep->p, 0xff, cap * 8, ctx_.stream));
      v.kv = nullptr;
      v.kv_dtype = 0;
      v.slot_rep = (int64_t*)im.slot_rep->p;
      v.rep_stride = 1;
    }
    if (n <= 0) break;
    BufPtr row_slot = dev_alloc(ctx_, (size_t)n * 4);
    BufPtr misc = dev_alloc_zero(ctx_, 32);  // u32 [0] has duplicates, [1] max count, [2..3] u64 total, [4] kv: a key equals the empty marker
    {
      KernelEvent ev(opt_.flags, ctx_.stream, use_kv ? "k_join_insert_kv" : "k_join_insert");
      if (use_kv) launch_join_insert_kv(v, (int32_t*)row_slot->p, (uint32_t*)misc->p, ctx_.stream);
      else launch_join_insert(v, (int32_t*)row_slot->p, (uint32_t*)misc->p, ctx_.stream);
    }
    if (deferred) {  // unique keys assumed; the caller checks the flags once the stream has been synchronised
      SQ_CUDA(cudaMemcpyAsync(hint_pinned_, misc->p, 24, cudaMemcpyDeviceToHost, ctx_.stream));
      ctx_.defer([row_slot, misc]() {});
      sealed_deferred_ = true;
      break;
    }
    uint32_t flags[6] = {0, 0, 0, 0, 0, 0};
    SQ_CUDA(cudaMemcpyAsync(flags, misc->p, sizeof(flags), cudaMemcpyDeviceToHost, ctx_.stream));
    SQ_CUDA(cudaStreamSynchronize(ctx_.stream));
    if (hint_pinned_) std::memcpy(hint_pinned_, flags, sizeof(flags));
    if (use_kv && flags[4]) {  // rare: rebuild in the other layout (the Bloom filter already holds a superset: harmless)
      use_kv = false;
      continue;
    }
    if (flags[0]) {
      // some key repeats: rows per slot, CSR ranges, ascending row ids per range
      v.unique = 0;
      im.slot_count = dev_alloc_zero(ctx_, cap * 4);
      im.slot_start = dev_alloc(ctx_, cap * 8);
      im.rows = dev_alloc(ctx_, (size_t)n * 8);
      v.slot_count = (uint32_t*)im.slot_count->p;
      v.slot_start = (uint64_t*)im.slot_start->p;
      v.rows = (int64_t*)im.rows->p;
      launch_join_count(v, (const int32_t*)row_slot->p, (uint32_t*)misc->p + 1, ctx_.stream);
      BufPtr scratch = dev_alloc(ctx_, scan_scratch_entries((int64_t)cap) * 8);
      launch_scan_u32_large(v.slot_count, (int64_t)cap, (unsigned long long*)v.slot_start, (unsigned long long*)misc->p + 1,
                            (unsigned long long*)scratch->p, ctx_.stream);
      uint32_t max_count = 0;
      SQ_CUDA(cudaMemcpyAsync(&max_count, (uint32_t*)misc->p + 1, 4, cudaMemcpyDeviceToHost, ctx_.stream));
      SQ_CUDA(cudaStreamSynchronize(ctx_.stream));
      im.max_count = max_count;
      if (max_count <= 64) {
        BufPtr fill = dev_alloc_zero(ctx_, cap * 4);
        launch_join_fill(v, (const int32_t*)row_slot->p, (uint32_t*)fill->p, ctx_.stream);
        launch_join_sort_ranges(v, ctx_.stream);
      } else {
        join_fill_sorted(v, (const int32_t*)row_slot->p, ctx_.stream);
      }
    }
    break;
  }
  if (join_type_ == SQLRS_JOIN_LEFT || join_type_ == SQLRS_JOIN_FULL) im.visited_left = dev_alloc_zero(ctx_, (size_t)bitmap_words(n) * 4 + 4);
}

void JoinOp::adopt_build(DBatch build, const JoinTableView& view, std::vector<BufPtr> keep) {
  Impl& im = *impl_;
  if (im.sealed || !im.left_batches.empty()) fail(SQLRS_ERR_INVALID_ARG, "hash_join: adopt_build on a join that already has a build side");
  im.sealed = true;
  im.left_rows = build.n;
  im.left_single = std::move(build);
  im.view = view;
  im.capacity = view.capacity;
  im.max_count = 1;
  im.adopted = std::move(keep);
}

std::string JoinOp::debug_probe_source(const std::vector<ColInfo>& probe_cols, const ExprCopy& probe_pred) const {
  return gen_input_decls(probe_cols) + gen_probe_program(probe_cols, right_keys_, probe_pred, match_keys()).src;
}

bool JoinOp::empty_build() const { return impl_->capacity == 0; }
int64_t JoinOp::build_rows() const { return impl_->left_rows; }
const JoinTableView& JoinOp::table_view() const { return impl_->view; }
const DBatch& JoinOp::build_side() const { return impl_->left_single; }

// build_batch, hash_join.rs:25-45: all left columns by (nullable) build index, all right columns by probe index
DBatch JoinOp::build_batch(const DBatch& right, const int64_t* li, bool li_nullable, const uint32_t* ri, int64_t m) {
  DBatch out;
  out.fields = out_fields_;
  out.n = m;
  const std::vector<bool>& needed = impl_->needed;
  auto pruned = [&](size_t k) {  // nobody above reads this column: a Null-typed placeholder keeps the positions
    if (needed.empty() || k >= needed.size() || needed[k]) return false;
    DCol c;
    c.dtype = SQLRS_DT_NULL;
    c.n = m;
    c.null_count = m;
    out.cols.push_back(c);
    return true;
  };
  size_t k = 0;
  for (const DCol& c : impl_->left_single.cols)
    if (!pruned(k++)) out.cols.push_back(gather_col_i64(ctx_, c, li, m, li_nullable));
  for (const DCol& c : right.cols)
    if (!pruned(k++)) out.cols.push_back(gather_col_u32(ctx_, c, ri, m));
  check_schema(out);
  return out;
}

// RecordBatch::try_new(schema, columns): column count / types / declared nullability must agree
void JoinOp::check_schema(DBatch& b) {
  if (b.cols.size() != out_fields_.size()) fail(SQLRS_ERR_ARROW, "number of columns must match number of fields in schema");
  for (size_t c = 0; c < b.cols.size(); c++) {
    if (!impl_->needed.empty() && c < impl_->needed.size() && !impl_->needed[c]) continue;  // pruned placeholder
    if (b.cols[c].dtype != out_fields_[c].dtype)
      fail(SQLRS_ERR_ARROW, std::string("column types must match schema types, expected ") + dtype_name(out_fields_[c].dtype) +
                                " but found " + dtype_name(b.cols[c].dtype));
    if (!out_fields_[c].nullable && null_count_of(ctx_, b.cols[c]) > 0)
      fail(SQLRS_ERR_ARROW, "Column '" + out_fields_[c].name + "' is declared as non-nullable but contains null values");
  }
}

// one probe batch, hash_join.rs:208-292
bool JoinOp::probe(const DBatch& right, DBatch* result) {
  seal();
  Impl& im = *impl_;
  if (im.capacity == 0) return false;  // empty build side: no left batch at all (:183-185)
  Trace tr("join.probe", ctx_.stream);
  ctx_.reap();
  const bool mk = opt_.match_mode == SQLRS_MATCH_HASH_AND_KEY;
  const bool keep_right = join_type_ == SQLRS_JOIN_RIGHT || join_type_ == SQLRS_JOIN_FULL;
  const int64_t n = right.n;
  if (n >= (1LL << 32)) fail(SQLRS_ERR_INVALID_ARG, "a probe batch may hold fewer than 2^32 rows (hash_join.rs:219)");
  // Fused probe (csrc/jit/joinprobe.cuh): ONE pass over the probe batch evaluates the Filter fused below the join and
  // the key expressions, tests the build side's Bloom filter and probes the table; it leaves the matched slot per row
  // and the output-row count per 2048-row chunk.  A scan over the chunk counts + k_join_probe_emit then produce the
  // (build row, probe row) pairs in the reference's order.
  const uint32_t* probe_keep = nullptr;  // Right/Full + non-equi filter: which probe rows passed the fused Filter
  BufPtr probe_keep_buf;
  int64_t total = 0;
  BufPtr li, ri;
  if (n > 0) {
    if (im.max_count >= (1u << 20)) fail(SQLRS_ERR_UNSUPPORTED, "join: more than 2^20 build rows share one key");
    Trace tr_a("  probe.fused count+scan+emit", ctx_.stream);
    std::vector<ColInfo> pcols = col_infos(right);
    const std::string sig = RowProgram(pcols).signature();
    auto kit = im.probe_kernels.find(sig);
    if (kit == im.probe_kernels.end()) {
      const std::string src = gen_input_decls(pcols) + gen_probe_program(pcols, right_keys_, im.probe_pred, mk).src;
      kit = im.probe_kernels.emplace(sig, jit_get("join_table+joinprobe", src, "sq_joinprobe_kernel")).first;
    }
    const int64_t chunks = div_up(n, kProbeChunk);
    BufPtr slot_of = dev_alloc(ctx_, (size_t)n * 4);
    BufPtr counts = dev_alloc(ctx_, (size_t)chunks * 4);
    BufPtr offsets = dev_alloc(ctx_, (size_t)chunks * 8 + 8);
    BufPtr err = dev_alloc_zero(ctx_, 8);
    {
      std::vector<const void*> in_blob(2 * std::max<size_t>(right.cols.size(), 1), nullptr);
      const size_t nc = std::max<size_t>(right.cols.size(), 1);
      for (size_t c = 0; c < right.cols.size(); c++) {
        in_blob[c] = right.cols[c].data;
        in_blob[nc + c] = right.cols[c].valid;
      }
      int64_t n_arg = n;
      JoinTableView jv = im.view;
      int keep = keep_right ? 1 : 0;
      void* slot_p = slot_of->p;
      void* counts_p = counts->p;
      void* err_p = err->p;
      void* args[] = {in_blob.data(), &n_arg, &jv, &keep, &slot_p, &counts_p, &err_p};
      // persistent grid: exactly the CTAs that are resident at once, so the chunk-stride loop has no ragged second wave
      const int per_sm = std::max(1, jit_max_blocks_per_sm(kit->second, 256, 0));
      const unsigned grid = (unsigned)std::min<int64_t>(chunks, (int64_t)device_sm_count(ctx_.device) * per_sm);
      KernelEvent ev(opt_.flags, ctx_.stream, "sq_joinprobe_kernel");
      jit_launch(kit->second, grid, 256, 0, ctx_.stream, args);
    }
    unsigned long long* total_d = (unsigned long long*)offsets->p + chunks;
    launch_scan_u32((const uint32_t*)counts->p, chunks, (unsigned long long*)offsets->p, total_d, ctx_.stream);
    struct {
      unsigned long long total;
      uint32_t err;
    } host = {0, 0};
    SQ_CUDA(cudaMemcpyAsync(&host.total, total_d, 8, cudaMemcpyDeviceToHost, ctx_.stream));
    SQ_CUDA(cudaMemcpyAsync(&host.err, err->p, 4, cudaMemcpyDeviceToHost, ctx_.stream));
    SQ_CUDA(cudaStreamSynchronize(ctx_.stream));
    if (host.err & 1u) fail(SQLRS_ERR_ARROW, std::string(arithmetic_error_text()) + " (join key)");
    total = (int64_t)host.total;
    if (total >= (1LL << 32)) fail(SQLRS_ERR_UNSUPPORTED, "join output of one probe batch exceeds 2^32 rows");
    li = dev_alloc(ctx_, (size_t)std::max<int64_t>(total, 1) * 8);
    ri = dev_alloc(ctx_, (size_t)std::max<int64_t>(total, 1) * 4);
    launch_join_probe_emit(im.view, (const int32_t*)slot_of->p, (const unsigned long long*)offsets->p, n, keep_right ? 1 : 0, (int64_t*)li->p,
                           (uint32_t*)ri->p, ctx_.stream);
    if (keep_right && !filter_.empty() && !im.probe_pred.empty()) {
      // apply_join_filter re-appends probe rows that lost all their matches — but only rows the fused Filter kept
      probe_keep_buf = dev_alloc(ctx_, (size_t)bitmap_words(n) * 4);
      launch_slot_keep_bitmap((const int32_t*)slot_of->p, n, (uint32_t*)probe_keep_buf->p, ctx_.stream);
      probe_keep = (const uint32_t*)probe_keep_buf->p;
    }
    ctx_.defer([slot_of, counts, offsets, err]() {});
  } else {
    li = dev_alloc(ctx_, 8);
    ri = dev_alloc(ctx_, 4);
  }

  if (!filter_.empty()) {  // apply_join_filter, :47-127
    DBatch inter = build_batch(right, (const int64_t*)li->p, keep_right, (const uint32_t*)ri->p, total);
    EvalResult mask = im.filter_prog->run(ctx_, inter, "join filter");
    int64_t kept = 0;
    BufPtr pos = compact_indices(ctx_, (const uint32_t*)mask.cols[0].data, total, &kept);
    int64_t extra = 0;
    BufPtr unvisited;
    if (keep_right && n > 0) {  // :73-121 — right rows that lost all their matches come back with a NULL left side
      BufPtr fr = dev_alloc(ctx_, (size_t)std::max<int64_t>(kept, 1) * 4);
      launch_gather_u32idx(4, ri->p, nullptr, (const uint32_t*)pos->p, kept, fr->p, nullptr, ctx_.stream);
      BufPtr visited = dev_alloc_zero(ctx_, (size_t)bitmap_words(n) * 4);
      launch_mark_bits_u32((const uint32_t*)fr->p, kept, (uint32_t*)visited->p, ctx_.stream);
      BufPtr inv = dev_alloc(ctx_, (size_t)bitmap_words(n) * 4);
      launch_bitmap_not((const uint32_t*)visited->p, n, (uint32_t*)inv->p, ctx_.stream, probe_keep);
      unvisited = compact_indices(ctx_, (const uint32_t*)inv->p, n, &extra);
    }
    const int64_t m = kept + extra;
    BufPtr fl = dev_alloc(ctx_, (size_t)std::max<int64_t>(m, 1) * 8), fr = dev_alloc(ctx_, (size_t)std::max<int64_t>(m, 1) * 4);
    launch_gather_u32idx(8, li->p, nullptr, (const uint32_t*)pos->p, kept, fl->p, nullptr, ctx_.stream);
    launch_gather_u32idx(4, ri->p, nullptr, (const uint32_t*)pos->p, kept, fr->p, nullptr, ctx_.stream);
    if (extra > 0) {
      SQ_CUDA(cudaMemsetAsync((int64_t*)fl->p + kept, 0xff, (size_t)extra * 8, ctx_.stream));  // -1 = NULL index
      SQ_CUDA(cudaMemcpyAsync((uint32_t*)fr->p + kept, unvisited->p, (size_t)extra * 4, cudaMemcpyDeviceToDevice, ctx_.stream));
    }
    li = fl;
    ri = fr;
    total = m;
  }
  if (im.visited_left && total > 0) launch_mark_bits_i64((const int64_t*)li->p, total, (uint32_t*)im.visited_left->p, ctx_.stream);  // :274-282
  *result = build_batch(right, (const int64_t*)li->p, keep_right, (const uint32_t*)ri->p, total);  // :284-291
  return true;
}

// Left/Full tail, hash_join.rs:296-322
bool JoinOp::finish(DBatch* result) {
  seal();
  Impl& im = *impl_;
  if (im.capacity == 0) return false;
  if (!(join_type_ == SQLRS_JOIN_LEFT || join_type_ == SQLRS_JOIN_FULL)) return false;
  const int64_t n = im.left_rows;
  int64_t m = 0;
  BufPtr idx;
  if (n > 0) {
    BufPtr inv = dev_alloc(ctx_, (size_t)bitmap_words(n) * 4);
    launch_bitmap_not((const uint32_t*)im.visited_left->p, n, (uint32_t*)inv->p, ctx_.stream, im.view.build_keep);
    idx = compact_indices(ctx_, (const uint32_t*)inv->p, n, &m);
  } else {
    idx = dev_alloc(ctx_, 4);
  }
  DBatch out;
  out.fields = out_fields_;
  out.n = m;
  for (size_t c = 0; c < im.left_single.cols.size(); c++) {
    if (!im.needed.empty() && c < im.needed.size() && !im.needed[c]) {
      DCol ph;
      ph.dtype = SQLRS_DT_NULL;
      ph.n = m;
      ph.null_count = m;
      out.cols.push_back(ph);
    } else {
      out.cols.push_back(gather_col_u32(ctx_, im.left_single.cols[c], (const uint32_t*)idx->p, m));
    }
  }
  for (size_t c = im.left_single.cols.size(); c < out_fields_.size(); c++) out.cols.push_back(null_col(ctx_, out_fields_[c].dtype, m));
  check_schema(out);
  *result = out;
  return true;
}

// ------------------------------------------------------------------ join chain (csrc/jit/joinchain.cuh)
namespace {
// rows per lane per trip of sq_joinchain_kernel (SQ_CUNROLL; tuning experiments: SQLRS_B200_CUNROLL)
int chain_unroll() {
  static const int u = [] {
    const char* e = std::getenv("SQLRS_B200_CUNROLL");
    return e ? std::min(16, std::max(1, atoi(e))) : 8;
  }();
  return u;
}
int64_t chain_trip_rows() { return 256LL * chain_unroll(); }  // SQ_CBLOCK x SQ_CUNROLL
}  // namespace

JoinChainOp::JoinChainOp(const Options& opt) : ctx_(opt), opt_(opt) {}
JoinChainOp::~JoinChainOp() {
  if (host_) cudaFreeHost(host_);
}

std::string JoinChainOp::debug_source(const std::vector<ColInfo>& build_cols, const std::vector<ColInfo>& probe_cols, const std::vector<ExprCopy>& right_keys1,
                                      const ExprCopy& probe_pred1, const ExprCopy& key2, int* key_dtype) {
  std::ostringstream s;
  s << "#define SQ_CUNROLL " << chain_unroll() << "\n";
  s << gen_input_decls(probe_cols) << gen_build_decls(build_cols);
  s << gen_probe_program(probe_cols, right_keys1, probe_pred1, true).src;
  RowProgram prog(build_cols, probe_cols);  // joined mode: join 1's output row
  Val k = prog.compile(key2, 1);
  if (k.dtype == SQLRS_DT_NULL || k.dtype == SQLRS_DT_UTF8) fail(SQLRS_ERR_UNSUPPORTED, "join chain: unsupported key type");
  const int raw = prog.emit_raw_bits(k);
  const int h = prog.emit_mix_hash({raw}, {k});  // the placement hash gen_probe_program gives join 2's probe side
  if (key_dtype) *key_dtype = k.dtype;
  bool uses_build = false;  // does join 2's key read a build-1 column?  (then it cannot be evaluated before join 1's probe resolved)
  for (const ExprNodeCopy& nd : key2) uses_build |= nd.op == SQLRS_OP_INPUT_REF && nd.index < (int)build_cols.size();
  s << "#define SQ_CHAIN_KEY_USES_BUILD " << (uses_build ? 1 : 0) << "\n";
  s << "struct SqChainKey { u64 h; u64 kb; u32 knull; };\n";
  s << "__device__ __forceinline__ void sq_chain_key(const SqIn& in, const SqInB& inb, i64 r, i64 b, SqChainKey& o, bool& e1) {\n  bool e0 = false;\n";
  s << prog.body_str();
  s << "  e1 |= e0;\n  o.h = v" << h << ";\n  o.kb = v" << raw << ";\n  o.knull = n" << k.id << " ? 0u : 1u;\n}\n";
  return s.str();
}

bool JoinChainOp::check_flags() {
  if (j1_deferred_) {  // join 1 was sealed without a synchronisation: repeated key / unrepresentable key / table full?
    j1_deferred_ = false;
    const uint32_t* g = host_->j1_flags;
    if (g[3]) fail(SQLRS_ERR_ARROW, std::string(arithmetic_error_text()) + " (join key)");
    if (g[0] || g[4] || g[5]) {
      hint_inserted_ = -1;  // the next run goes the synchronised way (which handles all of these)
      hint_j1_kept_ = -1;
      return false;
    }
  }
  const uint32_t* f = host_->flags;
  if (f[3]) fail(SQLRS_ERR_ARROW, std::string(arithmetic_error_text()) + " (join key)");
  if (f[0] || f[2]) {  // a repeated key needs the CSR lists, an unrepresentable key the slot_rep layout: not this path
    disabled_ = true;
    hint_inserted_ = -1;
    return false;
  }
  if (f[1]) {  // the table was too small for the estimate
    hint_inserted_ = -1;
    return false;
  }
  return true;
}

bool JoinChainOp::validate() {
  if (!pending_) return true;
  pending_ = false;
  if (!check_flags()) return false;
  hint_inserted_ = (int64_t)host_->inserted;
  hint_j1_kept_ = (int64_t)host_->j1_kept;
  return true;
}

bool JoinChainOp::run(JoinOp& j1, const DBatch& probe1, const ExprCopy& probe_pred1, const ExprCopy& key2, JoinOp& j2) {
  if (disabled_ || std::getenv("SQLRS_B200_NO_CHAIN")) return false;
  Trace tr("join.chain", ctx_.stream);
  ctx_.activate();
  ctx_.reap();
  const int64_t n = probe1.n;
  if (n <= 0 || n >= (1LL << 32)) return false;
  if (!host_) {
    SQ_CUDA(cudaHostAlloc((void**)&host_, sizeof(Host), cudaHostAllocDefault));
    std::memset(host_, 0, sizeof(Host));
  }
  // repeated run over tables of the same size: everything is sized from the previous run's numbers, nothing synchronises,
  // and validate() checks the flags afterwards
  const bool optimistic = hint_inserted_ >= 0 && hint_probe_rows_ == n && hint_build_rows_ == j1.build_rows() && !(opt_.flags & SQLRS_FLAG_TIMING);
  std::memset(host_->j1_flags, 0, sizeof(host_->j1_flags));
  j1.set_build_hint(optimistic ? hint_j1_kept_ : -1, host_->j1_flags);
  j1.seal();
  j1_deferred_ = j1.sealed_deferred();
  if (j1.empty_build()) return false;
  const JoinTableView& jt = j1.table_view();
  if (!jt.unique || jt.capacity == 0) return false;
  const DBatch& build1 = j1.build_side();
  std::vector<ColInfo> pcols = col_infos(probe1), bcols = col_infos(build1);
  const std::string sig = RowProgram(bcols, pcols).signature();
  auto kit = kernels_.find(sig);
  if (kit == kernels_.end()) {
    int key_dtype = 0;
    const std::string src = debug_source(bcols, pcols, j1.right_keys(), probe_pred1, key2, &key_dtype);
    kit = kernels_.emplace(sig, std::make_pair(jit_get("join_table+joinchain", src, "sq_joinchain_kernel"), key_dtype)).first;
  }
  JitKernel* kernel = kit->second.first;

  const int sms = device_sm_count(ctx_.device);
  const int per_sm = std::max(1, jit_max_blocks_per_sm(kernel, 256, 0));
  std::vector<const void*> in_blob(2 * std::max<size_t>(probe1.cols.size(), 1), nullptr), inb_blob(2 * std::max<size_t>(build1.cols.size(), 1), nullptr);
  for (size_t c = 0; c < probe1.cols.size(); c++) {
    in_blob[c] = probe1.cols[c].data;
    in_blob[std::max<size_t>(probe1.cols.size(), 1) + c] = probe1.cols[c].valid;
  }
  for (size_t c = 0; c < build1.cols.size(); c++) {
    inb_blob[c] = build1.cols[c].data;
    inb_blob[std::max<size_t>(build1.cols.size(), 1) + c] = build1.cols[c].valid;
  }
  struct ChainOut {
    uint64_t* kv;
    uint32_t* bloom;
    uint32_t capacity, bloom_mask;
    uint32_t* flags;
    unsigned long long* inserted;
  };
  BufPtr status = dev_alloc_zero(ctx_, 32);  // u32 flags[4] + u64 inserted
  auto launch = [&](ChainOut out, int64_t chunk_step) {
    int64_t n_arg = n, step = chunk_step;
    JoinTableView jv = jt;
    void* args[] = {in_blob.data(), inb_blob.data(), &n_arg, &jv, &out, &step};
    const int64_t trips = div_up(div_up(n, chain_trip_rows()), chunk_step);
    const unsigned grid = (unsigned)std::min<int64_t>(std::max<int64_t>(trips, 1), (int64_t)sms * per_sm);
    KernelEvent ev(opt_.flags, ctx_.stream, out.kv ? "sq_joinchain_kernel" : "sq_joinchain_kernel (sample count)");
    jit_launch(kernel, grid, 256, 0, ctx_.stream, args);
  };

  // ---- how many rows will join 1 yield?  Exact count of the previous run over tables of the same size, else a strided sample
  int64_t est = hint_inserted_;
  if (!optimistic) {
    const int64_t chunks = div_up(n, chain_trip_rows());
    const int64_t step = std::max<int64_t>(1, chunks / 4096);  // ~8 M sampled rows at most
    ChainOut cnt{nullptr, nullptr, 0, 0, (uint32_t*)status->p, (unsigned long long*)((uint32_t*)status->p + 4)};
    launch(cnt, step);
    SQ_CUDA(cudaMemcpyAsync(host_, status->p, 24, cudaMemcpyDeviceToHost, ctx_.stream));
    SQ_CUDA(cudaStreamSynchronize(ctx_.stream));
    if (host_->flags[3]) fail(SQLRS_ERR_ARROW, std::string(arithmetic_error_text()) + " (join key)");
    est = (int64_t)((double)host_->inserted * (double)step * 1.25) + 4096;
    SQ_CUDA(cudaMemsetAsync(status->p, 0, 32, ctx_.stream));
  }
  uint64_t cap = 1024;
  while (cap * kv_max_load_pct() < 100ULL * (uint64_t)est) cap <<= 1;
  if (cap > (1ULL << 30)) return false;
  cap_used_ = cap;
  BufPtr kv = dev_alloc(ctx_, cap * 16);
  SQ_CUDA(cudaMemsetAsync(kv->p, 0xff, cap * 16, ctx_.stream));
  const uint32_t bloom_words = join_bloom_words(est);
  BufPtr bloom = dev_alloc_zero(ctx_, (size_t)bloom_words * 4);
  ChainOut out{(uint64_t*)kv->p, (uint32_t*)bloom->p, (uint32_t)cap, bloom_words - 1, (uint32_t*)status->p, (unsigned long long*)((uint32_t*)status->p + 4)};
  launch(out, 1);
  SQ_CUDA(cudaMemcpyAsync(host_, status->p, 24, cudaMemcpyDeviceToHost, ctx_.stream));
  hint_probe_rows_ = n;
  hint_build_rows_ = j1.build_rows();
  int64_t inserted = est;
  if (optimistic) {
    pending_ = true;  // validated by the plan once the stream has been synchronised
  } else {
    SQ_CUDA(cudaStreamSynchronize(ctx_.stream));
    if (!check_flags()) return false;
    inserted = (int64_t)host_->inserted;
    hint_inserted_ = inserted;
    hint_j1_kept_ = (int64_t)host_->j1_kept;
  }
  ctx_.defer([status]() {});

  // ---- join 2's build side: a view of join 1's probe batch (row ids = its row numbers); build-1 columns are not available
  JoinTableView v{};
  v.capacity = (uint32_t)cap;
  v.kv = (uint64_t*)kv->p;
  v.kv_dtype = kit->second.second;
  v.slot_rep = (int64_t*)kv->p + 1;
  v.rep_stride = 2;
  v.n_build = n;
  v.n_keys = 1;
  v.match_keys = 1;
  v.bloom = (uint32_t*)bloom->p;
  v.bloom_mask = bloom_words - 1;
  v.unique = 1;
  v.n_inserted = inserted;
  DBatch vb;
  vb.fields = j1.out_fields();
  vb.n = n;
  for (size_t c = 0; c < build1.cols.size(); c++) {
    DCol ph;
    ph.dtype = SQLRS_DT_NULL;
    ph.n = n;
    ph.null_count = n;
    vb.cols.push_back(ph);
  }
  for (const DCol& c : probe1.cols) vb.cols.push_back(c);
  if (vb.cols.size() != vb.fields.size()) fail(SQLRS_ERR_ARROW, "number of columns must match number of fields in schema");
  j2.adopt_build(std::move(vb), v, {kv, bloom});
  return true;
}

}  // namespace sq
