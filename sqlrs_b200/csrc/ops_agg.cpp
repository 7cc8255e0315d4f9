// sqlrs_b200 — SimpleAggExecutor / HashAggExecutor on the GPU (fused with a Filter child when the
// plan has one).  Reference: src/executor/aggregate/{simple_agg.rs:27-65, hash_agg.rs:33-150,
// sum.rs:16-97, count.rs:10-29, min_max.rs:47-157}; kernels: csrc/jit/agg.cuh.
#include <algorithm>

#include "join.hpp"
#include "kernels_aot.hpp"
#include "ops.hpp"

namespace sq {

std::vector<AggSpec> copy_aggs(const sqlrs_agg_desc* aggs, int32_t n) {
  std::vector<AggSpec> out;
  if (n > 0 && !aggs) fail(SQLRS_ERR_INVALID_ARG, "aggs is NULL");
  for (int32_t k = 0; k < n; k++) {
    AggSpec a;
    a.func = aggs[k].func;
    a.distinct = aggs[k].distinct;
    a.return_dtype = aggs[k].return_dtype;
    a.arg = copy_expr(&aggs[k].arg);
    a.name = aggs[k].name ? aggs[k].name : "";
    out.push_back(a);
  }
  return out;
}

namespace {

enum WordOp { W_ADD_U64, W_ADD_F64, W_MIN_I64, W_MAX_I64, W_COUNT_EPOCH };

struct WordPlan {
  int op;
};
struct AggPlan {
  int func, out_dtype, arg_dtype;
  int value_word = -1;   // main accumulator word
  int nvalid_word = -1;  // number of non-NULL inputs (only when the argument can be NULL)
  bool f64_sortable = false;
  bool utf8_packed = false;  // MIN / MAX over a Utf8 column: the word holds rank << 32 | string pool id
};

constexpr uint64_t kEpochShift = 40;
constexpr uint64_t kEpochMask = (1ULL << kEpochShift) - 1;

uint64_t word_identity(int op) {
  switch (op) {
    case W_MIN_I64: return 0x7fffffffffffffffULL;
    case W_MAX_I64: return 0x8000000000000000ULL;
  }
  return 0;
}

inline uint32_t next_pow2(uint64_t v) {
  uint64_t p = 1024;
  while (p < v) p <<= 1;
  if (p > (1ULL << 31)) fail(SQLRS_ERR_INTERNAL, "group table would exceed 2^31 slots");
  return (uint32_t)p;
}

struct SqInBlob {
  std::vector<const void*> blob;
  SqInBlob(const DBatch& b, int64_t row_start) {
    size_t n = b.cols.size() ? b.cols.size() : 1;
    blob.assign(2 * n, nullptr);
    for (size_t c = 0; c < b.cols.size(); c++) {
      const DCol& col = b.cols[c];
      if (col.data) {
        if (col.dtype == SQLRS_DT_BOOL) blob[c] = (const uint32_t*)col.data + (row_start >> 5);
        else blob[c] = (const uint8_t*)col.data + (size_t)row_start * dtype_width(col.dtype);
      }
      if (col.valid) blob[n + c] = col.valid + (row_start >> 5);
    }
  }
  void* ptr() { return blob.data(); }
};

// CUDA events around the dominant scan kernel on the operator's own stream (SQLRS_FLAG_TIMING)
struct ScanTimer {
  cudaStream_t stream;
  bool enabled;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  ScanTimer(cudaStream_t s, bool on) : stream(s), enabled(on) {
    if (!enabled) return;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, stream);
  }
  void stop() {
    if (enabled) cudaEventRecord(e1, stream);
  }
  double elapsed_ms() {  // call after the stream has been synchronised
    if (!enabled) return 0.0;
    float ms = 0.f;
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
  }
  void release(cudaEvent_t* a, cudaEvent_t* b) {  // hands the events over (a launch whose end nobody waits for here)
    *a = e0;
    *b = e1;
    e0 = e1 = nullptr;
  }
  ~ScanTimer() {
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  }
};

}  // namespace

// COUNT(DISTINCT x) / SUM(DISTINCT x) (reference: DistinctCountAccumulator count.rs:31-58, DistinctSumAccumulator
// sum.rs:99-132 — a HashSet<ScalarValue> per group).  On the GPU the set becomes one more level of grouping:
//   dedup   = GROUP BY (keys..., x) without aggregates            (the set elements, true equality, NULL is a value)
//   second  = GROUP BY keys over dedup's output: COUNT(1) — the reference's set counts a NULL element too — or SUM(x)
//   plain   = the operator's non-DISTINCT aggregates, as usual
// All three emit their groups in first-appearance order of the keys over the same input, so their rows line up and the
// result is the positional zip [keys, aggregates in declared order].
struct AggOp::Distinct {
  std::unique_ptr<AggOp> plain;
  struct Item {
    size_t agg_index;
    std::unique_ptr<AggOp> dedup;
  };
  std::vector<Item> items;
  bool seen_batch = false;
};

struct HostGroups {
  uint32_t n = 0;
  std::vector<uint64_t> hash, min_row, keys, acc;  // keys [K][n], acc [W][n]
  std::vector<uint32_t> knull;
};

struct AggOp::Compiled {
  JitKernel *small = nullptr, *merge = nullptr, *global = nullptr, *fixkeys = nullptr, *medium = nullptr;
  int mslots = 0, munroll = 4, medium_grid = 0;  // sq_agg_medium: groups per CTA, rows per thread per trip
  size_t medium_smem = 0;
  bool medium_ok = false;
  std::vector<AggPlan> aggs;
  std::vector<WordPlan> words;
  std::vector<int> key_dtypes;
  std::vector<bool> key_decl_null;  // may the key be NULL in some batch of this schema?
  std::vector<int> tile_cols;       // fused probe->aggregate: probe columns the TMA variant stages in shared memory
  std::vector<int> utf8_minmax_cols;  // input columns under MIN / MAX(Utf8): pushed as rank << 32 | id (AggOp::push)
  int block = 128, slots = 8, unroll = 4, min_ctas = 1;
  size_t small_smem = 0;
  int small_grid = 0;
  bool small_ok = true;
};

struct AggOp::Table {
  uint32_t capacity = 0;
  int n_keys = 0, n_acc = 0;
  BufPtr state, hash, min_row, keys, knull, acc, new_slots, counters;
  std::vector<uint64_t> identities;
  TableView view() const {
    TableView v;
    v.state = (uint32_t*)state->p;
    v.hash = (uint64_t*)hash->p;
    v.min_row = (uint64_t*)min_row->p;
    v.keys = (uint64_t*)keys->p;
    v.knull = (uint32_t*)knull->p;
    v.acc = (uint64_t*)acc->p;
    v.new_slots = (uint32_t*)new_slots->p;
    v.counters = (uint32_t*)counters->p;
    v.capacity = capacity;
    return v;
  }
};

AggOp::AggOp(std::vector<AggSpec> aggs, std::vector<ExprCopy> group_by, std::vector<std::string> group_names, bool simple,
             ExprCopy fused_predicate, const Options& opt)
    : ctx_(opt), opt_(opt), aggs_(std::move(aggs)), group_by_(std::move(group_by)), group_names_(std::move(group_names)),
      simple_(simple), predicate_(std::move(fused_predicate)) {
  if (group_by_.size() > 16) fail(SQLRS_ERR_UNSUPPORTED, "more than 16 group-by keys");
  bool any_distinct = false;
  for (const AggSpec& a : aggs_) {
    if (a.func < SQLRS_AGG_COUNT || a.func > SQLRS_AGG_MAX) fail(SQLRS_ERR_INVALID_ARG, "unknown aggregate function");
    // create_accumulator (aggregate/mod.rs:27-49): only Count and Sum have DISTINCT accumulators, Min/Max ignore the flag
    any_distinct |= a.distinct && (a.func == SQLRS_AGG_COUNT || a.func == SQLRS_AGG_SUM);
  }
  if (any_distinct) {
    if (group_by_.size() >= 16) fail(SQLRS_ERR_UNSUPPORTED, "DISTINCT aggregates with 16 group-by keys");
    distinct_ = std::make_unique<Distinct>();
    Options sub = opt;
    sub.stream = ctx_.stream;  // every sub-operator works on this operator's stream
    sub.device_id = ctx_.device;
    std::vector<AggSpec> plain_aggs;
    for (size_t j = 0; j < aggs_.size(); j++) {
      const AggSpec& a = aggs_[j];
      if (a.distinct && (a.func == SQLRS_AGG_COUNT || a.func == SQLRS_AGG_SUM)) {
        Options dopt = sub;
        dopt.match_mode = SQLRS_MATCH_HASH_AND_KEY;  // HashSet<ScalarValue>: real equality, whatever the group identity mode
        std::vector<ExprCopy> keys = group_by_;
        keys.push_back(a.arg);
        std::vector<std::string> names = group_names_;
        names.resize(group_by_.size());
        names.push_back("distinct_arg");
        distinct_->items.push_back({j, std::make_unique<AggOp>(std::vector<AggSpec>{}, keys, names, false, predicate_, dopt)});
      } else {
        plain_aggs.push_back(a);
      }
    }
    // (an ungrouped aggregate whose aggregates are all DISTINCT needs no plain operator: it yields exactly one row)
    if (!(plain_aggs.empty() && group_by_.empty()))
      distinct_->plain = std::make_unique<AggOp>(plain_aggs, group_by_, group_names_, simple_, predicate_, sub);
  }
}
AggOp::~AggOp() {
  if (pinned_) cudaFreeHost(pinned_);
  if (pend_e0_) cudaEventDestroy(pend_e0_);
  if (pend_e1_) cudaEventDestroy(pend_e1_);
}

void AggOp::resolve_pending_timer() {
  if (!pend_e1_) return;
  float ms = 0.f;
  cudaEventSynchronize(pend_e1_);
  cudaEventElapsedTime(&ms, pend_e0_, pend_e1_);
  scan_kernel_ms_ += ms;
  scan_kernel_launches_ += 1;
  cudaEventDestroy(pend_e0_);
  cudaEventDestroy(pend_e1_);
  pend_e0_ = pend_e1_ = nullptr;
}

// ------------------------------------------------------------------ code generation
AggOp::Compiled& AggOp::compiled_for(const DBatch& batch) {
  std::vector<ColInfo> cols = col_infos(batch);
  std::string sig = RowProgram(cols).signature();
  auto it = cache_.find(sig);
  if (it != cache_.end()) return *it->second;

  auto comp = std::make_unique<Compiled>();
  std::string src = generate(cols, *comp);
  comp->small = comp->small_ok ? jit_get("agg_table+agg", src, "sq_agg_small") : nullptr;
  comp->merge = jit_get("agg_table+agg", src, "sq_agg_merge");
  comp->global = jit_get("agg_table+agg", src, "sq_agg_global");
  comp->fixkeys = jit_get("agg_table+agg", src, "sq_agg_fixkeys");
  if (comp->medium_ok) {
    comp->medium = jit_get("agg_table+agg", src, "sq_agg_medium");
    int per_sm = jit_max_blocks_per_sm(comp->medium, 256, comp->medium_smem);
    if (per_sm < 1) comp->medium_ok = false;
    comp->medium_grid = device_sm_count(ctx_.device) * std::max(per_sm, 1);
  }
  if (comp->small_ok) {
    int per_sm = jit_max_blocks_per_sm(comp->small, comp->block, comp->small_smem);
    if (per_sm < 1) comp->small_ok = false;
    comp->small_grid = device_sm_count(ctx_.device) * std::max(per_sm, 1);
  }
  if (!key_dtypes_.empty() && key_dtypes_ != comp->key_dtypes) fail(SQLRS_ERR_ARROW, "group key types changed between batches");
  if (!cache_.empty()) {
    const Compiled& first = *cache_.begin()->second;
    bool same = first.words.size() == comp->words.size();
    for (size_t w = 0; same && w < first.words.size(); w++) same = first.words[w].op == comp->words[w].op;
    if (!same) fail(SQLRS_ERR_ARROW, "batch schema changed between batches (a column declared non-nullable contains nulls?)");
  }
  Compiled& ref = *comp;
  cache_[sig] = std::move(comp);
  return ref;
}

std::string AggOp::debug_source(const std::vector<ColInfo>& cols) {
  Compiled c;
  return generate(cols, c);
}

std::string AggOp::debug_join_source(const std::vector<ColInfo>& build_cols, const std::vector<ColInfo>& probe_cols,
                                     const std::vector<ExprCopy>& right_keys, const ExprCopy& probe_pred, const ExprCopy& join_filter) {
  JoinGen jg;
  jg.build_cols = build_cols;
  jg.right_keys = right_keys;
  jg.probe_pred = probe_pred;
  jg.join_filter = join_filter;
  jg.jmatch = opt_.match_mode == SQLRS_MATCH_HASH_AND_KEY;
  Compiled c;
  return generate(probe_cols, c, &jg);
}

// the row program + accumulator glue for csrc/jit/agg.cuh (pure: no device access)
std::string AggOp::generate(const std::vector<ColInfo>& cols, Compiled& comp_ref, const JoinGen* jg) {
  Compiled* comp = &comp_ref;
  // plain: one program over the batch row.  joined (jg): the program addresses the joined row of an inner hash
  // join — build columns gathered at the matched build row, probe columns at the scan row; its "fused predicate"
  // is the join's non-equi filter (the probe-side Filter and the join keys live in the stage-1 program below)
  std::unique_ptr<RowProgram> prog_holder(jg ? new RowProgram(jg->build_cols, cols) : new RowProgram(cols));
  RowProgram& prog = *prog_holder;
  const ExprCopy& stage_pred = jg ? jg->join_filter : predicate_;
  const bool fused = !stage_pred.empty() || jg != nullptr;
  std::string pass = "true";
  if (!stage_pred.empty()) {
    Val p = prog.compile(stage_pred, jg ? 1 : 0);
    if (p.dtype != SQLRS_DT_BOOL) fail(SQLRS_ERR_INTERNAL, "filter executor expected evaluate boolean array");
    pass = "(n" + std::to_string(p.id) + " && v" + std::to_string(p.id) + ")";
  }
  const int ec = fused ? 1 : 0;
  // hash_agg.rs:63-73: aggregate arguments first, then the group keys
  std::vector<Val> args;
  for (const AggSpec& a : aggs_) args.push_back(prog.compile(a.arg, ec));
  std::vector<Val> keys;
  for (const ExprCopy& g : group_by_) keys.push_back(prog.compile(g, ec));
  for (const Val& k : keys) {
    comp->key_dtypes.push_back(k.dtype);
    comp->key_decl_null.push_back(k.decl_null || k.maybe_null || k.always_null);
  }

  // accumulator words
  std::ostringstream upd;  // body of sq_acc_update
  for (size_t j = 0; j < aggs_.size(); j++) {
    const AggSpec& a = aggs_[j];
    AggPlan p;
    p.func = a.func;
    p.arg_dtype = args[j].dtype;
    Val v = args[j];
    auto add_word = [&](int op) {
      comp->words.push_back(WordPlan{op});
      return (int)comp->words.size() - 1;
    };
    switch (a.func) {
      case SQLRS_AGG_COUNT: {
        p.out_dtype = SQLRS_DT_INT64;
        p.value_word = add_word(opt_.count_mode == SQLRS_COUNT_REFERENCE_OVERWRITE ? W_COUNT_EPOCH : W_ADD_U64);
        upd << "  a[" << p.value_word << " * stride] += " << (v.dtype == SQLRS_DT_NULL ? std::string("0ULL") : "(o.an" + std::to_string(j) + " ? 1ULL : 0ULL)")
            << ";\n";
        break;
      }
      case SQLRS_AGG_SUM: {
        // sum.rs:54 casts to the return type, arrow `sum` skips NULLs, ints wrap (release build)
        if (v.dtype == SQLRS_DT_UTF8 || a.return_dtype == SQLRS_DT_UTF8) fail(SQLRS_ERR_UNSUPPORTED, "unsupported sum type: Utf8");
        Val c = prog.cast(v, a.return_dtype);
        if (!is_numeric(c.dtype)) fail(SQLRS_ERR_UNSUPPORTED, std::string("unsupported sum type: ") + dtype_name(c.dtype));
        if (c.dtype == SQLRS_DT_INT32)
          fail(SQLRS_ERR_UNSUPPORTED, "not expected Int32 and Int32 for sum");  // sum.rs:84 unimplemented!
        p.out_dtype = c.dtype;
        std::string vn = "o.a" + std::to_string(j), nn = "o.an" + std::to_string(j);
        if (c.dtype == SQLRS_DT_FLOAT64) {
          p.value_word = add_word(W_ADD_F64);
          upd << "  if (" << nn << ") { double* p = (double*)(a + " << p.value_word << " * stride); *p = __dadd_rn(*p, " << vn << "); }\n";
        } else {
          p.value_word = add_word(W_ADD_U64);
          upd << "  if (" << nn << ") a[" << p.value_word << " * stride] += (u64)" << vn << ";\n";
        }
        if (c.decl_null) {
          p.nvalid_word = add_word(W_ADD_U64);
          upd << "  a[" << p.nvalid_word << " * stride] += " << nn << " ? 1ULL : 0ULL;\n";
        }
        args[j] = c;
        break;
      }
      case SQLRS_AGG_MIN:
      case SQLRS_AGG_MAX: {
        if (v.dtype == SQLRS_DT_UTF8) {
          // min_string / max_string (min_max.rs:12-19,47-65): the column arrives as rank << 32 | id (AggOp::push re-expresses
          // ids through the string pool's current byte-wise ranks), so the Int64 order of the words is the strings' order
          if (jg || a.arg.size() != 1 || a.arg[0].op != SQLRS_OP_INPUT_REF)
            fail(SQLRS_ERR_UNSUPPORTED, "min/max over a Utf8 EXPRESSION (only a plain column, outside fused joins) is not supported by the CUDA backend");
          if (std::find(comp->utf8_minmax_cols.begin(), comp->utf8_minmax_cols.end(), a.arg[0].index) == comp->utf8_minmax_cols.end())
            comp->utf8_minmax_cols.push_back(a.arg[0].index);
        } else if (!is_numeric(v.dtype)) {
          fail(SQLRS_ERR_UNSUPPORTED, std::string("unsupported min/max type: ") + dtype_name(v.dtype));
        }
        if (v.dtype != a.return_dtype)
          fail(SQLRS_ERR_UNSUPPORTED, std::string("unsupported min_max scalar type: ") + dtype_name(a.return_dtype));
        p.out_dtype = v.dtype;
        const bool is_min = a.func == SQLRS_AGG_MIN;
        p.value_word = add_word(is_min ? W_MIN_I64 : W_MAX_I64);
        p.f64_sortable = v.dtype == SQLRS_DT_FLOAT64;
        p.utf8_packed = v.dtype == SQLRS_DT_UTF8;
        std::string vn = "o.a" + std::to_string(j), nn = "o.an" + std::to_string(j);
        std::string as_i64 = p.f64_sortable ? "sq_f64_sortable(" + vn + ")" : "(i64)" + vn;
        upd << "  if (" << nn << ") { i64* p = (i64*)(a + " << p.value_word << " * stride); const i64 x = " << as_i64 << "; if (x "
            << (is_min ? "<" : ">") << " *p) *p = x; }\n";
        if (v.decl_null) {
          p.nvalid_word = add_word(W_ADD_U64);
          upd << "  a[" << p.nvalid_word << " * stride] += " << nn << " ? 1ULL : 0ULL;\n";
        }
        break;
      }
    }
    comp->aggs.push_back(p);
  }
  const int K = (int)keys.size(), W = (int)comp->words.size();
  if (W > 48) fail(SQLRS_ERR_UNSUPPORTED, "too many aggregate accumulators for one operator (max 48 words)");
  // a column under MIN / MAX(Utf8) is pushed in its packed form: nothing else of the operator may read its ids
  for (int c : comp->utf8_minmax_cols) {
    auto reads = [&](const ExprCopy& e) {
      for (const ExprNodeCopy& nd : e)
        if (nd.op == SQLRS_OP_INPUT_REF && nd.index == c) return true;
      return false;
    };
    bool other = reads(stage_pred);
    for (const ExprCopy& g : group_by_) other = other || reads(g);
    for (const AggSpec& a : aggs_)
      if (!(a.func == SQLRS_AGG_MIN || a.func == SQLRS_AGG_MAX || a.func == SQLRS_AGG_COUNT)) other = other || reads(a.arg);
    if (other) fail(SQLRS_ERR_UNSUPPORTED, "a Utf8 column under MIN / MAX that is also a group key, predicate input or SUM argument is not supported by the CUDA backend");
  }
  std::vector<int> raw_ids;
  for (const Val& k : keys) raw_ids.push_back(prog.emit_raw_bits(k));
  // group identity: the reference's row hash when it IS the identity (hash-only, quirk K2); with key
  // comparison any hash places the group, so a cheaper multiply-xorshift mix of the key bits is used
  const int hash_id = opt_.match_mode == SQLRS_MATCH_HASH_AND_KEY ? prog.emit_mix_hash(raw_ids, keys) : prog.emit_row_hash(keys);

  // launch shape of sq_agg_small: S slots, T threads, private accumulators in shared memory
  comp->slots = K == 0 ? 1 : 8;
  // rows per thread per trip: 8 keeps ~64 8-byte loads in flight per thread for narrow scans (Q1': +4 % over 4);
  // wide scans stay at 4 to bound registers
  comp->unroll = 4;
  comp->block = 128;
  auto smem_for = [&](int T) {
    return (size_t)(W + 1) * comp->slots * T * 8 + (size_t)8 * comp->slots * (8 + 8 + 8 * std::max(K, 1) + 4 + 4 + 4) + 16;  // + SQ_TSLOTS = 8 S slot-table entries (tag, hash, keys, null mask, state, group)
  };
  if (smem_for(256) <= 100 * 1024) comp->block = 256;
  // tuning overrides (experiments only): SQLRS_B200_AGG_BLOCK / _UNROLL / _SLOTS
  if (const char* e = std::getenv("SQLRS_B200_AGG_BLOCK")) comp->block = std::max(32, atoi(e) / 32 * 32);
  if (const char* e = std::getenv("SQLRS_B200_AGG_UNROLL")) comp->unroll = std::max(1, atoi(e));
  if (const char* e = std::getenv("SQLRS_B200_AGG_SLOTS"))
    if (K > 0) comp->slots = std::max(1, atoi(e));
  while (comp->block > 32 && smem_for(comp->block) > 200 * 1024) comp->block /= 2;
  comp->small_ok = smem_for(comp->block) <= 200 * 1024;
  comp->small_smem = smem_for(comp->block);
  // rows per thread per trip (measured on Q1' SF10, profiles/): 6 keeps ~48 8-byte loads in flight per thread and is
  // the best for narrow scans with the cheap placement hash (0.673 ms vs 0.710 at 4, 0.811 at 8); with the
  // reference's ahash identity the longer program is fastest at 4.  The register budget is handed to the compiler
  // as __launch_bounds__(block, min CTAs) so that a wider program cannot silently drop a resident CTA.
  if (!std::getenv("SQLRS_B200_AGG_UNROLL"))
    comp->unroll = (comp->block == 128 && prog.n_loaded_columns() <= 8 && opt_.match_mode == SQLRS_MATCH_HASH_AND_KEY) ? 6 : 4;
  const int smem_ctas = (int)std::max<size_t>(1, (200 * 1024) / std::max<size_t>(comp->small_smem, 1));
  const int reg_ctas = 65536 / (comp->block * (comp->unroll >= 6 ? 170 : 128));
  comp->min_ctas = std::max(1, std::min(smem_ctas, reg_ctas));
  // sq_agg_medium: one accumulator copy per CTA for up to M groups; M = the largest power of two whose shared
  // memory (accumulators + 2M-entry slot table) stays <= 72 KB, i.e. three 256-thread CTAs per SM
  auto medium_smem_for = [&](int M) { return (size_t)(W + 1) * M * 8 + (size_t)2 * M * (8 + 8 + 8 * std::max(K, 1) + 12) + 16; };
  comp->mslots = 0;
  for (int M = 2048; M >= 64; M /= 2)
    if (medium_smem_for(M) <= 72 * 1024) {
      comp->mslots = M;
      break;
    }
  comp->medium_ok = K > 0 && comp->mslots > comp->slots;
  comp->medium_smem = comp->medium_ok ? medium_smem_for(comp->mslots) : 0;
  comp->munroll = prog.n_loaded_columns() <= 8 ? 4 : 2;

  std::ostringstream s;
  s << gen_input_decls(cols);
  if (jg) {
    // build-side inputs + the stage-1 program: fused probe-side Filter, join keys, their row hash (the same
    // create_hashes the build side was hashed with)
    const size_t nb = jg->build_cols.size() ? jg->build_cols.size() : 1;
    s << "struct SqInB { const void* col[" << nb << "]; const u32* val[" << nb << "]; };\n";
    s << "#define SQ_LDB_I64(c, b) sq_ldg_i64(inb.col[c], b)\n#define SQ_LDB_I32(c, b) sq_ldg_i32(inb.col[c], b)\n";
    s << "#define SQ_LDB_F64(c, b) sq_ldg_f64(inb.col[c], b)\n#define SQ_LDB_BOOL(c, b) sq_ld_bit(inb.col[c], b)\n";
    s << "#define SQ_VALIDB(c, b) sq_ld_bit(inb.val[c], b)\n";
    ProbeProgram pp = gen_probe_program(cols, jg->right_keys, jg->probe_pred, jg->jmatch);
    s << pp.src;
    comp->tile_cols = pp.tile_cols;
  }
  s << "#define SQ_NKEYS " << K << "\n#define SQ_NACC " << W << "\n";
  s << "#define SQ_MATCH_KEYS " << (opt_.match_mode == SQLRS_MATCH_HASH_AND_KEY ? 1 : 0) << "\n";
  s << "#define SQ_MSLOTS " << std::max(comp->mslots, 64) << "\n#define SQ_MUNROLL " << comp->munroll << "\n";
  s << "#define SQ_MINCTAS " << comp->min_ctas << "\n";
  s << "#define SQ_BLOCK " << comp->block << "\n#define SQ_SLOTS " << comp->slots << "\n#define SQ_UNROLL " << comp->unroll << "\n";
  s << "struct SqRow {\n  bool pass; u64 h; u64 kb[" << std::max(K, 1) << "]; u32 knull;\n";
  for (size_t j = 0; j < aggs_.size(); j++) s << "  " << ctype_of(args[j].dtype) << " a" << j << "; bool an" << j << ";\n";
  s << "};\n";
  if (jg) s << "__device__ __forceinline__ void sq_row(const SqIn& in, const SqInB& inb, i64 r, i64 b, SqRow& o, bool& e1) {\n  bool e0 = false;\n";
  else s << "__device__ __forceinline__ void sq_row(const SqIn& in, i64 r, SqRow& o, bool& e0, bool& e1) {\n";
  s << prog.body_str();
  if (jg) s << "  e1 |= e0;\n";
  s << "  o.pass = " << pass << ";\n  o.h = v" << hash_id << ";\n";
  if (K == 0) s << "  o.kb[0] = 0ULL;\n";
  std::string knull = "0u";
  for (int k = 0; k < K; k++) {
    s << "  o.kb[" << k << "] = v" << raw_ids[k] << ";\n";
    knull += " | (n" + std::to_string(keys[k].id) + " ? 0u : " + std::to_string(1u << k) + "u)";
  }
  s << "  o.knull = " << knull << ";\n";
  for (size_t j = 0; j < aggs_.size(); j++)
    s << "  o.a" << j << " = v" << args[j].id << "; o.an" << j << " = n" << args[j].id << ";\n";
  s << "}\n";
  s << "__device__ __forceinline__ u64 sq_acc_identity(int w) {\n  switch (w) {\n";
  for (int w = 0; w < W; w++)
    if (word_identity(comp->words[w].op) != 0) s << "    case " << w << ": return 0x" << std::hex << word_identity(comp->words[w].op) << std::dec << "ULL;\n";
  s << "  }\n  return 0ULL;\n}\n";
  s << "__device__ __forceinline__ void sq_acc_update(u64* a, int stride, const SqRow& o) {\n" << upd.str() << "}\n";
  s << "__device__ __forceinline__ u64 sq_acc_reduce(int w, u64 x, u64 y) {\n  switch (w) {\n";
  for (int w = 0; w < W; w++) {
    s << "    case " << w << ": ";
    switch (comp->words[w].op) {
      case W_ADD_U64:
      case W_COUNT_EPOCH: s << "return x + y;\n"; break;
      case W_ADD_F64: s << "return (u64)__double_as_longlong(__dadd_rn(__longlong_as_double((i64)x), __longlong_as_double((i64)y)));\n"; break;
      case W_MIN_I64: s << "return (u64)(((i64)y < (i64)x) ? (i64)y : (i64)x);\n"; break;
      case W_MAX_I64: s << "return (u64)(((i64)y > (i64)x) ? (i64)y : (i64)x);\n"; break;
    }
  }
  s << "  }\n  return x;\n}\n";
  // the same fold into a CTA-shared copy (sq_agg_medium): plain atomics, the batch epoch is applied at the flush
  s << "__device__ __forceinline__ void sq_acc_merge_shared(u64* p, int w, u64 x) {\n  switch (w) {\n";
  for (int w = 0; w < W; w++) {
    s << "    case " << w << ": ";
    switch (comp->words[w].op) {
      case W_ADD_U64:
      case W_COUNT_EPOCH: s << "if (x) atomicAdd(p, x); break;\n"; break;
      case W_ADD_F64: s << "atomicAdd((double*)p, __longlong_as_double((i64)x)); break;\n"; break;
      case W_MIN_I64: s << "atomicMin((i64*)p, (i64)x); break;\n"; break;
      case W_MAX_I64: s << "atomicMax((i64*)p, (i64)x); break;\n"; break;
    }
  }
  s << "  }\n}\n";
  s << "__device__ __forceinline__ void sq_acc_merge_global(u64* p, int w, u64 x, i64 batch_no) {\n  switch (w) {\n";
  for (int w = 0; w < W; w++) {
    s << "    case " << w << ": ";
    switch (comp->words[w].op) {
      case W_ADD_U64: s << "if (x) atomicAdd(p, x); break;\n"; break;
      case W_ADD_F64: s << "atomicAdd((double*)p, __longlong_as_double((i64)x)); break;\n"; break;
      case W_MIN_I64: s << "atomicMin((i64*)p, (i64)x); break;\n"; break;
      case W_MAX_I64: s << "atomicMax((i64*)p, (i64)x); break;\n"; break;
      case W_COUNT_EPOCH:
        // CountAccumulator::update_batch ASSIGNS (count.rs:22, quirk K1): the value is the count of the
        // last batch that touched the group.  Word = (batch epoch << 40) | count within that batch.
        s << "{ const u64 ep = (u64)(batch_no + 1) << " << kEpochShift << "; u64 old = *p;\n"
          << "      for (;;) { const u64 nv = ((old >> " << kEpochShift << ") == (u64)(batch_no + 1)) ? old + x : (ep | x);\n"
          << "        const u64 prev = atomicCAS(p, old, nv); if (prev == old) break; old = prev; } break; }\n";
        break;
    }
  }
  s << "  }\n}\n";

  return s.str();
}

std::string AggOp::describe() const { return last_path_; }

// ------------------------------------------------------------------ table
// counters: [0] groups, [1] new-slot list length, [2] status bits, [3] error flag (divide by zero)
void AggOp::init_table_contents(Table& t) {
  const Compiled& c = *cache_.begin()->second;
  const size_t cap = t.capacity;
  SQ_CUDA(cudaMemsetAsync(t.state->p, 0, cap * 4, ctx_.stream));
  SQ_CUDA(cudaMemsetAsync(t.min_row->p, 0xff, cap * 8, ctx_.stream));
  SQ_CUDA(cudaMemsetAsync(t.counters->p, 0, 16, ctx_.stream));
  for (int w = 0; w < t.n_acc; w++) {
    uint64_t ident = word_identity(c.words[w].op);
    if (ident == 0) SQ_CUDA(cudaMemsetAsync((uint64_t*)t.acc->p + (size_t)w * cap, 0, cap * 8, ctx_.stream));
    else launch_fill_u64((uint64_t*)t.acc->p + (size_t)w * cap, (int64_t)cap, ident, ctx_.stream);
  }
}

void AggOp::ensure_table(uint32_t min_capacity) {
  if (table_ && table_->capacity >= min_capacity) return;
  grow_table(min_capacity);
}

std::unique_ptr<AggOp::Table> AggOp::new_table(uint32_t capacity) {
  const Compiled& c = *cache_.begin()->second;
  auto t = std::make_unique<Table>();
  t->capacity = next_pow2(capacity);
  t->n_keys = (int)c.key_dtypes.size();
  t->n_acc = (int)c.words.size();
  const size_t cap = t->capacity;
  t->state = dev_alloc(ctx_, cap * 4);
  t->hash = dev_alloc(ctx_, cap * 8);
  t->min_row = dev_alloc(ctx_, cap * 8);
  t->keys = dev_alloc_zero(ctx_, cap * 8 * std::max(t->n_keys, 1));
  t->knull = dev_alloc_zero(ctx_, cap * 4);
  t->acc = dev_alloc(ctx_, cap * 8 * std::max(t->n_acc, 1));
  t->new_slots = dev_alloc(ctx_, cap * 4);
  t->counters = dev_alloc(ctx_, 16);
  init_table_contents(*t);
  return t;
}

void AggOp::grow_table(uint32_t min_capacity) {
  slot_list_complete_ = false;
  const Compiled& c = *cache_.begin()->second;
  auto t = std::make_unique<Table>();
  t->capacity = next_pow2(min_capacity);
  t->n_keys = (int)c.key_dtypes.size();
  t->n_acc = (int)c.words.size();
  const size_t cap = t->capacity;
  t->state = dev_alloc(ctx_, cap * 4);
  t->hash = dev_alloc(ctx_, cap * 8);
  t->min_row = dev_alloc(ctx_, cap * 8);
  t->keys = dev_alloc_zero(ctx_, cap * 8 * std::max(t->n_keys, 1));
  t->knull = dev_alloc_zero(ctx_, cap * 4);
  t->acc = dev_alloc(ctx_, cap * 8 * std::max(t->n_acc, 1));
  t->new_slots = dev_alloc(ctx_, cap * 4);
  t->counters = dev_alloc(ctx_, 16);
  init_table_contents(*t);
  if (table_) {
    launch_table_rehash(table_->view(), t->view(), t->n_keys, t->n_acc, ctx_.stream);
    // carry the counters (the pending new-slot list is flushed before any growth)
    SQ_CUDA(cudaMemcpyAsync(t->counters->p, table_->counters->p, 4, cudaMemcpyDeviceToDevice, ctx_.stream));
    SQ_CUDA(cudaMemcpyAsync((uint32_t*)t->counters->p + 2, (uint32_t*)table_->counters->p + 2, 8, cudaMemcpyDeviceToDevice, ctx_.stream));
  }
  table_ = std::move(t);
}

// forget all groups but keep the compiled kernels and the device buffers (repeated plan runs)
void AggOp::reset() {
  ctx_.activate();
  if (distinct_) {
    if (distinct_->plain) distinct_->plain->reset();
    for (auto& it : distinct_->items) it.dedup->reset();
    distinct_->seen_batch = false;
  }
  slot_list_complete_ = false;
  slot_list_pending_ = false;
  hint_sized_ = false;
  tier_pending_ = false;
  resolve_pending_timer();  // (into the totals that are zeroed below)
  if (table_) init_table_contents(*table_);
  rows_seen_ = 0;
  batches_seen_ = 0;
  seen_batch_ = false;
  // the tier that fits the previous run's group count (a plan is usually re-run over similar data): the escalation small ->
  // medium -> global costs two wasted passes over the first batch otherwise
  level_ = 0;
  if (!cache_.empty() && groups_hint_ > 0) {
    const Compiled& c = *cache_.begin()->second;
    if (groups_hint_ > (uint32_t)c.slots) level_ = 1;
    if (groups_hint_ > (uint32_t)std::max(c.mslots, c.slots)) level_ = 2;
  }
  groups_known_ = 0;
  groups_bound_ = 0;
  counters_stale_ = false;
  scan_kernel_ms_ = 0;
  scan_kernel_launches_ = 0;
}

// one D2H of the 4 counters; raises what the kernels flagged
void AggOp::read_counters(uint32_t* out4) {
  SQ_CUDA(cudaMemcpyAsync(out4, table_->counters->p, 16, cudaMemcpyDeviceToHost, ctx_.stream));
  SQ_CUDA(cudaStreamSynchronize(ctx_.stream));
  groups_known_ = out4[0];
  groups_bound_ = out4[0];
  counters_stale_ = false;
  groups_hint_ = out4[0];
  if (slot_list_pending_) {  // deferred push_join: the launch appended every new group to new_slots
    slot_list_pending_ = false;
    slot_list_complete_ = out4[1] == out4[0];
    SQ_CUDA(cudaMemsetAsync((uint32_t*)table_->counters->p + 1, 0, 4, ctx_.stream));
  }
  if (tier_pending_) {  // a deferred sq_agg_small launch: did its per-CTA slots hold every group?
    tier_pending_ = false;
    if (out4[2] & 1u) {
      defer_disabled_ = true;  // this operator's input has more groups than the small tier holds: never defer again
      defer_tier_check_ = false;
      SQ_CUDA(cudaMemsetAsync((uint32_t*)table_->counters->p + 2, 0, 4, ctx_.stream));
      throw RetrySizingError{};
    }
  }
  if (hint_sized_) {
    hint_sized_ = false;
    if (out4[2] & 2u) {  // more groups than the previous run's count allowed for: the plan re-runs with exact sizing
      groups_hint_ = 0;
      throw RetrySizingError{};
    }
  }
  if (out4[2] & 2u) fail(SQLRS_ERR_INTERNAL, "group table overflow (internal sizing error)");
  if (out4[3] & 1u) {
    SQ_CUDA(cudaMemsetAsync((uint32_t*)table_->counters->p + 3, 0, 4, ctx_.stream));
    fail(SQLRS_ERR_ARROW, std::string(arithmetic_error_text()) + " (aggregate argument)");
  }
}

// CTA-partial scratch of sq_agg_small / sq_agg_medium, kept across batches
void AggOp::ensure_partial_scratch(size_t entries, int K, size_t W) {
  if (part_entries_ >= entries) return;
  p_state_ = dev_alloc(ctx_, entries * 4);
  p_hash_ = dev_alloc(ctx_, entries * 8);
  p_min_ = dev_alloc(ctx_, entries * 8);
  p_keys_ = dev_alloc(ctx_, entries * 8 * std::max(K, 1));
  p_knull_ = dev_alloc(ctx_, entries * 4);
  p_acc_ = dev_alloc(ctx_, entries * 8 * std::max<size_t>(W, 1));
  part_entries_ = entries;
}

// ------------------------------------------------------------------ push
void AggOp::push(const DBatch& batch_in) {
  if (distinct_) {
    distinct_->seen_batch = true;
    seen_batch_ = true;
    const DBatch& batch = batch_in;
    if (distinct_->plain) distinct_->plain->push(batch);
    for (auto& it : distinct_->items) it.dedup->push(batch);
    last_path_ = "DISTINCT: GROUP BY (keys, argument) dedup + second-level aggregate" +
                 (distinct_->plain ? "; plain aggregates: " + distinct_->plain->describe() : std::string());
    return;
  }
  Trace tr("agg.push", ctx_.stream);
  slot_list_complete_ = false;
  ctx_.activate();
  ctx_.reap();
  if (tier_pending_) settle();  // a second batch after a deferred first one: its tier is decided now
  Compiled& c0 = compiled_for(batch_in);
  // MIN / MAX over Utf8 columns: re-express what has been accumulated so far and this batch's ids in the string pool's
  // CURRENT byte-wise ranks (the pool may have grown since the last batch), as rank << 32 | id
  DBatch packed_batch;
  if (!c0.utf8_minmax_cols.empty()) {
    const std::vector<int32_t> ranks = StringPool::instance().ranks();
    BufPtr rank_table = dev_alloc(ctx_, std::max<size_t>(ranks.size(), 1) * 4);
    if (!ranks.empty()) SQ_CUDA(cudaMemcpyAsync(rank_table->p, ranks.data(), ranks.size() * 4, cudaMemcpyHostToDevice, ctx_.stream));
    SQ_CUDA(cudaStreamSynchronize(ctx_.stream));  // `ranks` is a local vector
    if (table_)
      for (size_t j = 0; j < c0.aggs.size(); j++)
        if (c0.aggs[j].utf8_packed)
          launch_str_rerank((uint64_t*)table_->acc->p + (size_t)c0.aggs[j].value_word * table_->capacity, (const uint32_t*)table_->state->p, table_->capacity,
                            (const int32_t*)rank_table->p, word_identity(c0.words[c0.aggs[j].value_word].op), ctx_.stream);
    packed_batch = batch_in;
    for (int col : c0.utf8_minmax_cols) {
      if (col < 0 || col >= (int)packed_batch.cols.size()) fail(SQLRS_ERR_INTERNAL, "InputRef index out of bounds");
      DCol& src = packed_batch.cols[(size_t)col];
      if (src.dtype != SQLRS_DT_UTF8) continue;
      DCol dst = make_col(ctx_, SQLRS_DT_UTF8, src.n, false);
      launch_str_pack((const int64_t*)src.data, src.n, (const int32_t*)rank_table->p, (int64_t*)col_data(dst), ctx_.stream);
      dst.valid = src.valid;
      dst.keep_valid = src.keep_valid;
      dst.null_count = src.null_count;
      src = dst;
    }
    ctx_.defer([rank_table]() {});
  }
  const DBatch& batch = c0.utf8_minmax_cols.empty() ? batch_in : packed_batch;
  Compiled& c = c0;
  if (key_dtypes_.empty()) key_dtypes_ = c.key_dtypes;
  seen_batch_ = true;
  const int64_t n = batch.n;
  const int64_t batch_no = batches_seen_++;
  const int64_t row_base = rows_seen_;
  rows_seen_ += n;
  if (n == 0) return;
  if (n >= (1LL << 32)) fail(SQLRS_ERR_INVALID_ARG, "a batch may hold fewer than 2^32 rows (reference index width, hash_agg.rs:109)");
  const int K = (int)c.key_dtypes.size();
  const size_t W = c.words.size();
  uint32_t hc[4] = {0, 0, 0, 0};
  // key fix-up for the groups that appeared since the last flush (hash-only identity: the reference
  // reports the keys of a group's FIRST row), then reset the list.  Must run before the table is
  // re-hashed (slot numbers change) and at the end of every batch (the rows are gone afterwards).
  auto flush_new_slots = [&]() {
    const uint32_t n_new = hc[1];
    if (n_new == 0 || !table_) return;
    if (K > 0 && opt_.match_mode == SQLRS_MATCH_HASH_ONLY) {
      SqInBlob in_all(batch, 0);
      int64_t n_all = n, rb_all = row_base;
      TableView tv_all = table_->view();
      uint32_t nn = n_new;
      void* fargs[] = {in_all.ptr(), &n_all, &rb_all, &tv_all, &nn};
      jit_launch(c.fixkeys, (unsigned)div_up(n_new, 128), 128, 0, ctx_.stream, fargs);
    }
    SQ_CUDA(cudaMemsetAsync((uint32_t*)table_->counters->p + 1, 0, 4, ctx_.stream));
    hc[1] = 0;
  };
  // capacity for `extra` more groups; only syncs when the host-side bound says the table might not hold them
  auto reserve = [&](uint64_t extra) {
    if (!table_) {
      ensure_table((uint32_t)std::min<uint64_t>(2ULL * extra + 1024, 1ULL << 31));
    } else if (2ULL * (groups_bound_ + extra) + 1024 > table_->capacity) {
      read_counters(hc);  // exact count instead of the bound
      flush_new_slots();
      ensure_table((uint32_t)std::min<uint64_t>(2ULL * (groups_known_ + extra) + 1024, 1ULL << 31));
    }
    groups_bound_ += extra;
  };

  bool done = false;
  if (level_ == 0 && !c.small_ok) level_ = 1;
  if (level_ == 0) {
    const int grid = (int)std::min<int64_t>(c.small_grid, std::max<int64_t>(1, div_up(n, (int64_t)c.block * c.unroll)));
    const size_t entries = (size_t)grid * c.slots;
    reserve(entries);
    ensure_partial_scratch(entries, K, W);
    struct {
      void *state, *hash, *min_row, *keys, *knull, *acc;
    } part = {p_state_->p, p_hash_->p, p_min_->p, p_keys_->p, p_knull_->p, p_acc_->p};
    SqInBlob in(batch, 0);
    int64_t n_arg = n, rb = row_base, bn = batch_no;
    void* status = (uint32_t*)table_->counters->p + 2;
    void* errp = (uint32_t*)table_->counters->p + 3;
    TableView tv = table_->view();
    int n_entries = (int)entries;
    void* args_small[] = {in.ptr(), &n_arg, &rb, &part, &status, &errp};
    ScanTimer timer(ctx_.stream, (opt_.flags & SQLRS_FLAG_TIMING) != 0);
    {
      KernelEvent ev(opt_.flags, ctx_.stream, "sq_agg_small");
      jit_launch(c.small, (unsigned)grid, (unsigned)c.block, c.small_smem, ctx_.stream, args_small);
    }
    timer.stop();
    void* args_merge[] = {&part, &n_entries, &tv, &bn, &status};
    jit_launch(c.merge, (unsigned)div_up(n_entries, 128), 128, 0, ctx_.stream, args_merge);
    if (defer_tier_check_ && batch_no == 0 && (K == 0 || opt_.match_mode == SQLRS_MATCH_HASH_AND_KEY) && c.utf8_minmax_cols.empty()) {
      // partial/final split: nothing synchronises here; the overflow bit is looked at by the next counter read, or travels
      // in the header of the packed partial buffer (see set_defer_tier_check)
      tier_pending_ = true;
      counters_stale_ = true;
      resolve_pending_timer();
      if (timer.enabled) timer.release(&pend_e0_, &pend_e1_);
      SQ_CUDA(cudaMemsetAsync((uint32_t*)table_->counters->p + 1, 0, 4, ctx_.stream));  // (what flush_new_slots does)
      last_path_ = "sq_agg_small (private shared-memory accumulators, " + std::to_string(c.slots) + " slots x " + std::to_string(c.block) +
                   " threads, grid " + std::to_string(grid) + ") + sq_agg_merge, tier check deferred";
      return;
    }
    read_counters(hc);
    scan_kernel_ms_ += timer.elapsed_ms();
    scan_kernel_launches_ += timer.enabled ? 1 : 0;
    if (hc[2] & 1u) {
      level_ = 1;  // more groups than the private-accumulator path holds: this and later batches go one level up
      SQ_CUDA(cudaMemsetAsync((uint32_t*)table_->counters->p + 2, 0, 4, ctx_.stream));
    } else {
      done = true;
      last_path_ = "sq_agg_small (private shared-memory accumulators, " + std::to_string(c.slots) + " slots x " + std::to_string(c.block) +
                   " threads, grid " + std::to_string(grid) + ") + sq_agg_merge";
    }
  }
  if (!done && level_ == 1 && !c.medium_ok) level_ = 2;
  if (!done && level_ == 1) {
    const int grid = (int)std::min<int64_t>(c.medium_grid, std::max<int64_t>(1, div_up(n, (int64_t)256 * c.munroll)));
    const size_t entries = (size_t)grid * c.mslots;
    reserve(entries);
    ensure_partial_scratch(entries, K, W);
    struct {
      void *state, *hash, *min_row, *keys, *knull, *acc;
    } part = {p_state_->p, p_hash_->p, p_min_->p, p_keys_->p, p_knull_->p, p_acc_->p};
    SqInBlob in(batch, 0);
    int64_t n_arg = n, rb = row_base, bn = batch_no;
    void* status = (uint32_t*)table_->counters->p + 2;
    void* errp = (uint32_t*)table_->counters->p + 3;
    TableView tv = table_->view();
    int n_entries = (int)entries;
    void* args[] = {in.ptr(), &n_arg, &rb, &part, &status, &errp};
    ScanTimer timer(ctx_.stream, (opt_.flags & SQLRS_FLAG_TIMING) != 0);
    {
      KernelEvent ev(opt_.flags, ctx_.stream, "sq_agg_medium");
      jit_launch(c.medium, (unsigned)grid, 256, c.medium_smem, ctx_.stream, args);
    }
    timer.stop();
    void* args_merge[] = {&part, &n_entries, &tv, &bn, &status};
    jit_launch(c.merge, (unsigned)div_up(n_entries, 128), 128, 0, ctx_.stream, args_merge);
    read_counters(hc);
    scan_kernel_ms_ += timer.elapsed_ms();
    scan_kernel_launches_ += timer.enabled ? 1 : 0;
    if (hc[2] & 4u) {
      level_ = 2;  // more groups than one CTA's shared-memory table holds
      SQ_CUDA(cudaMemsetAsync((uint32_t*)table_->counters->p + 2, 0, 4, ctx_.stream));
    } else {
      done = true;
      last_path_ = "sq_agg_medium (shared-memory atomics, " + std::to_string(c.mslots) + " groups per CTA, grid " + std::to_string(grid) + ")";
    }
  }
  if (!done) {
    level_ = 2;
    // The table stays dense (see sq_agg_global): it starts at 64 k slots and accepts new groups up to 3/4 of its capacity; rows
    // of further new groups come back in an overflow list, the table grows x4 (re-hash) and the kernel runs again over the list.
    const int64_t chunk = 1LL << 24;
    const int sms = device_sm_count(ctx_.device);
    if (!table_) ensure_table(1u << 16);
    BufPtr lists[2] = {dev_alloc(ctx_, (size_t)std::min(chunk, n) * 4), nullptr};
    BufPtr ov_count = dev_alloc_zero(ctx_, 4);
    int grows = 0;
    for (int64_t start = 0; start < n; start += chunk) {
      const int64_t len = std::min(chunk, n - start);
      SqInBlob in(batch, start);
      const uint32_t* redo = nullptr;
      uint32_t n_redo = 0;
      int cur = 0;
      for (;;) {
        int64_t n_arg = len, rb = row_base + start, bn = batch_no;
        void* status = (uint32_t*)table_->counters->p + 2;
        void* errp = (uint32_t*)table_->counters->p + 3;
        TableView tv = table_->view();
        uint32_t limit = table_->capacity / 4 * 3;
        void* ov_rows = lists[cur]->p;
        void* ov_cnt = ov_count->p;
        void* args[] = {in.ptr(), &n_arg, &rb, &tv, &bn, &status, &errp, &redo, &n_redo, &ov_rows, &ov_cnt, &limit};
        const int64_t items = redo ? (int64_t)n_redo : len;
        unsigned grid = (unsigned)std::min<int64_t>(std::max<int64_t>(div_up(items, 256 * 2), 1), (int64_t)sms * 8);  // 256 threads x SQ_GUNROLL (2) rows per trip
        ScanTimer timer(ctx_.stream, (opt_.flags & SQLRS_FLAG_TIMING) != 0);
        {
          KernelEvent ev(opt_.flags, ctx_.stream, "sq_agg_global");
          jit_launch(c.global, grid, 256, 0, ctx_.stream, args);
        }
        timer.stop();
        uint32_t overflowed = 0;
        SQ_CUDA(cudaMemcpyAsync(&overflowed, ov_count->p, 4, cudaMemcpyDeviceToHost, ctx_.stream));
        read_counters(hc);  // synchronises: the group count, the status bits and `overflowed`
        scan_kernel_ms_ += timer.elapsed_ms();
        scan_kernel_launches_ += timer.enabled && !redo ? 1 : 0;
        if (std::getenv("SQLRS_B200_DEBUG_CHUNKS"))
          fprintf(stderr, "[sqlrs] sq_agg_global chunk @%lld: %lld items, %.3f ms, groups %u, overflowed %u, capacity %u\n", (long long)start, (long long)items,
                  timer.elapsed_ms(), hc[0], overflowed, table_->capacity);
        if (overflowed == 0) break;
        flush_new_slots();  // slot numbers change under the re-hash
        grow_table(table_->capacity * 4);
        grows++;
        SQ_CUDA(cudaMemsetAsync(ov_count->p, 0, 4, ctx_.stream));
        redo = (const uint32_t*)lists[cur]->p;
        n_redo = overflowed;
        cur ^= 1;
        if (!lists[cur]) lists[cur] = dev_alloc(ctx_, (size_t)std::min(chunk, n) * 4);
      }
    }
    last_path_ = "sq_agg_global (dense open-addressed table in HBM, capacity " + std::to_string(table_->capacity) + ", grown " + std::to_string(grows) + "x)";
  }
  flush_new_slots();
}

// ------------------------------------------------------------------ fused probe -> aggregate
void AggOp::settle() {
  if (table_ && counters_stale_) {
    uint32_t hc[4];
    read_counters(hc);
  }
}

void AggOp::push_join(const DBatch& probe, JoinOp& join, const ExprCopy& probe_pred, bool defer) {
  if (distinct_) fail(SQLRS_ERR_INTERNAL, "push_join: DISTINCT aggregates take the unfused path");
  Trace tr("agg.push_join", ctx_.stream);
  ctx_.activate();
  ctx_.reap();
  settle();  // an earlier deferred batch
  if (join.join_type() != SQLRS_JOIN_INNER || opt_.match_mode != SQLRS_MATCH_HASH_AND_KEY)
    fail(SQLRS_ERR_INTERNAL, "push_join: only inner joins with key comparison are fused");
  join.seal();
  if (join.empty_build()) return;  // the join yields no batch at all
  const DBatch& build = join.build_side();
  std::vector<ColInfo> pcols = col_infos(probe), bcols = col_infos(build);
  const std::string sig = RowProgram(bcols, pcols).signature();
  JoinGen jg;
  jg.build_cols = bcols;
  jg.right_keys = join.right_keys();
  jg.probe_pred = probe_pred;
  jg.join_filter = join.join_filter();
  jg.jmatch = join.match_keys();
  auto kit = join_kernels_.find(sig);
  if (kit == join_kernels_.end()) {
    auto comp = std::make_unique<Compiled>();
    std::string src = generate(pcols, *comp, &jg);
    JoinKernels k;
    k.generic = jit_get("agg_table+join_table+joinagg", src, "sq_joinagg_kernel");
    k.tile_cols = comp->tile_cols;
    // opt-in (SQLRS_B200_TMA=1): measured 2.1x SLOWER than the register-staged kernel on Q3' (profiles/r01j_*: 632 vs 300 us
    // at SF10, 7.2 vs 2.6 ms at SF100) — 64 KB of tile ring per CTA cuts the resident warps from 40 to 24 per SM and the
    // per-tile CTA barrier couples every warp to the slowest phase-B warp, while the register-staged loads of 40 warps
    // already cover the DRAM latency
    if (!k.tile_cols.empty() && std::getenv("SQLRS_B200_TMA")) k.tma = jit_get("agg_table+join_table+joinagg", src, "sq_joinagg_tma_kernel");
    if (!cache_.empty()) {
      const Compiled& first = *cache_.begin()->second;
      bool same = first.words.size() == comp->words.size() && first.key_dtypes == comp->key_dtypes;
      for (size_t w = 0; same && w < first.words.size(); w++) same = first.words[w].op == comp->words[w].op;
      if (!same) fail(SQLRS_ERR_ARROW, "batch schema changed between batches (a column declared non-nullable contains nulls?)");
    } else {
      cache_["join|" + sig] = std::move(comp);  // the accumulator layout of this operator
    }
    kit = join_kernels_.emplace(sig, k).first;
  }
  const Compiled& c = *cache_.begin()->second;
  if (key_dtypes_.empty()) key_dtypes_ = c.key_dtypes;
  seen_batch_ = true;
  const int64_t n = probe.n;
  const int64_t batch_no = batches_seen_++;
  const int64_t row_base = rows_seen_;
  rows_seen_ += n;
  if (n == 0) return;
  if (n >= (1LL << 32)) fail(SQLRS_ERR_INVALID_ARG, "a probe batch may hold fewer than 2^32 rows (hash_join.rs:219)");
  // first-appearance ordinals of the joined stream are (global probe row << 20 | match ordinal) in one u64
  if ((uint64_t)(row_base + n) >= (1ULL << 44)) fail(SQLRS_ERR_UNSUPPORTED, "fused probe + aggregate: global probe row numbers must stay below 2^44");

  // The number of joined rows (and groups) is unknown before the probe: aggregate into a batch-local table sized
  // optimistically; a full table discards it and retries 4x larger, so a batch counts all-or-nothing.
  const JoinTableView& jt = join.table_view();
  // guess: about as many groups as build rows in the table (x2 head-room in the open-addressed table); a repeated run
  // knows better: the previous run's group count (+25 %)
  const bool main_empty_before = !table_ || (groups_known_ == 0 && !counters_stale_ && groups_bound_ == 0);
  const bool use_hint = defer && groups_hint_ > 0 && main_empty_before && !(opt_.flags & SQLRS_FLAG_TIMING);
  uint64_t cap = std::max<uint64_t>(2ULL * (uint64_t)(jt.n_inserted > 0 ? jt.n_inserted : jt.n_build), 1ULL << 16);
  if (use_hint) cap = std::max<uint64_t>(2ULL * ((uint64_t)groups_hint_ + groups_hint_ / 4), 1ULL << 16);
  std::unique_ptr<Table> local;
  uint32_t hc[4] = {0, 0, 0, 0};
  bool first_try = true, tma_used = false;
  for (;;) {
    if (cap > (1ULL << 31)) fail(SQLRS_ERR_INTERNAL, "group table would exceed 2^31 slots");
    {
      Trace tr_t("  joinagg.new_table", ctx_.stream);
      // repeated plan runs: the operator's (re-initialised, still empty) table of the previous run is the batch-local table
      // (a hint-sized run wants exactly the hinted capacity: a smaller table keeps more of itself in L2)
      const bool reuse = table_ && groups_known_ == 0 && !counters_stale_ && groups_bound_ == 0 && first_try &&
                         (use_hint ? table_->capacity == next_pow2(cap) : table_->capacity >= next_pow2(cap));
      if (reuse) local = std::move(table_);
      else local = new_table((uint32_t)cap);
      first_try = false;
    }
    Trace tr_k("  joinagg.kernel", ctx_.stream);
    SqInBlob in(probe, 0), inb(build, 0);
    int64_t rb = row_base, bn = batch_no;
    void* status = (uint32_t*)local->counters->p + 2;
    void* errp = (uint32_t*)local->counters->p + 3;
    TableView tv = local->view();
    JoinTableView jv = jt;
    const int sms = device_sm_count(ctx_.device);
    const JoinKernels& jk = kit->second;
    // TMA variant: whole 2048-row slices of the probe program's columns are staged in shared memory by bulk-async
    // copies, double buffered — needs 16-byte aligned column buffers (sliced Arrow arrays may not be)
    const int kTmaTileRows = tma_shape().tile_rows, kTmaStages = tma_shape().stages, kTmaConsumerWarps = tma_shape().consumers;
    bool use_tma = jk.tma != nullptr && n >= kTmaTileRows;
    for (int c : jk.tile_cols)
      use_tma = use_tma && c < (int)probe.cols.size() && probe.cols[(size_t)c].data && ((uintptr_t)probe.cols[(size_t)c].data % 16 == 0);
    ScanTimer timer(ctx_.stream, (opt_.flags & SQLRS_FLAG_TIMING) != 0);
    int64_t done_rows = 0;
    if (use_tma) {
      int64_t n_tiles = n / kTmaTileRows;
      const size_t smem = (size_t)kTmaStages * jk.tile_cols.size() * kTmaTileRows * 8;
      const int tblock = 32 * (kTmaConsumerWarps + 1);  // one producer warp + the consumer warps
      const int per_sm = std::max(1, jit_max_blocks_per_sm(jk.tma, tblock, smem));
      unsigned grid = (unsigned)std::min<int64_t>(n_tiles, (int64_t)sms * per_sm);
      void* args[] = {in.ptr(), inb.ptr(), &n_tiles, &rb, &jv, &tv, &bn, &status, &errp};
      KernelEvent ev(opt_.flags, ctx_.stream, "sq_joinagg_tma_kernel");
      jit_launch(jk.tma, grid, (unsigned)tblock, smem, ctx_.stream, args);
      done_rows = n_tiles * kTmaTileRows;
    }
    if (done_rows < n) {  // everything, or the ragged tail behind the last full tile
      SqInBlob in_tail(probe, done_rows);
      int64_t n_arg = n - done_rows, rb_tail = row_base + done_rows;
      // persistent grid: exactly the CTAs that are resident at once (a ragged second wave cost ~25 % at 5 CTAs / SM)
      const int per_sm = std::max(1, jit_max_blocks_per_sm(jk.generic, 256, 0));
      unsigned grid = (unsigned)std::min<int64_t>(div_up(n_arg, 2048), (int64_t)sms * per_sm);  // 256 threads x SQ_JUNROLL (8) rows per trip
      void* args[] = {in_tail.ptr(), inb.ptr(), &n_arg, &rb_tail, &jv, &tv, &bn, &status, &errp};
      KernelEvent ev(opt_.flags, ctx_.stream, "sq_joinagg_kernel");
      jit_launch(jk.generic, grid, 256, 0, ctx_.stream, args);
    }
    timer.stop();
    tma_used = use_tma;
    if (use_hint) {  // nothing synchronises: the table becomes the operator's, its counters are read by the next consumer
      table_ = std::move(local);
      counters_stale_ = true;
      hint_sized_ = true;
      slot_list_pending_ = true;
      groups_bound_ = table_->capacity;
      last_path_ = "sq_joinagg_kernel (fused probe + aggregate, table sized by the previous run's group count: capacity " +
                   std::to_string(table_->capacity) + ", unsynchronised)";
      return;
    }
    SQ_CUDA(cudaMemcpyAsync(hc, local->counters->p, 16, cudaMemcpyDeviceToHost, ctx_.stream));
    SQ_CUDA(cudaStreamSynchronize(ctx_.stream));
    scan_kernel_ms_ += timer.elapsed_ms();
    scan_kernel_launches_ += timer.enabled ? 1 : 0;
    if (hc[3] & 1u) fail(SQLRS_ERR_ARROW, std::string(arithmetic_error_text()) + " (aggregate argument)");
    if (!(hc[2] & 2u)) break;
    cap *= 4;
  }
  last_path_ = std::string(tma_used ? "sq_joinagg_tma_kernel (TMA-staged probe tiles, " : "sq_joinagg_kernel (") +
               "fused probe + aggregate, batch-local table capacity " + std::to_string(local->capacity) + ")";
  const bool main_empty = !table_ || (groups_known_ == 0 && !counters_stale_ && groups_bound_ == 0);
  if (main_empty) {
    table_ = std::move(local);  // first batch: its table IS the operator's table
    SQ_CUDA(cudaMemsetAsync((uint32_t*)table_->counters->p + 1, 0, 4, ctx_.stream));
    groups_hint_ = hc[0];
    groups_known_ = hc[0];
    groups_bound_ = hc[0];
    counters_stale_ = false;
    slot_list_complete_ = hc[1] == hc[0];  // every group of this launch was appended to new_slots (the list itself stays in place)
    return;
  }
  if (hc[0] == 0) return;
  // later batches: fold the batch-local groups into the operator's table (packed rows, device to device)
  const int words = 3 + local->n_keys + local->n_acc;
  BufPtr packed = dev_alloc(ctx_, (size_t)(hc[0] + 1) * words * 8);
  SQ_CUDA(cudaMemsetAsync(packed->p, 0, (size_t)words * 8, ctx_.stream));
  launch_table_pack(local->view(), local->n_keys, local->n_acc, (uint64_t*)packed->p, hc[0], ctx_.stream);
  merge_partials_device((const uint64_t*)packed->p, 1, hc[0], true);
}

// ------------------------------------------------------------------ finish
// the n groups packed in first-appearance order; a table straight out of one fused probe+aggregate launch still has the
// complete list of its slots, and its order keys are (probe row << 20 | match ordinal) < rows_seen_ << 20
void AggOp::pack_sorted(int K, int W, uint32_t n, uint64_t* dst) {
  if (slot_list_complete_) {
    int bits = 21;
    for (int64_t r = rows_seen_; r > 0; r >>= 1) bits++;
    table_pack_sorted(table_->view(), K, W, n, dst, ctx_.stream, (const uint32_t*)table_->new_slots->p, std::min(bits, 64));
  } else {
    table_pack_sorted(table_->view(), K, W, n, dst, ctx_.stream);
  }
}

// the group table -> host, through ONE packed buffer and one synchronisation
void AggOp::build_output(std::vector<Field>* fields, HostGroups* g) {
  if (!seen_batch_) fail(SQLRS_ERR_INTERNAL, "called `Option::unwrap()` on a `None` value (no input batch)");
  ctx_.activate();
  const Compiled& c = *cache_.begin()->second;
  const int K = (int)c.key_dtypes.size(), W = (int)c.words.size();
  g->n = 0;
  if (table_ && counters_stale_ && groups_bound_ > 4096) {  // a device-side merge skipped its counter read
    uint32_t hc[4];
    read_counters(hc);
  }
  if (table_ && (counters_stale_ ? groups_bound_ > 0 : groups_known_ > 0)) {
    // exact count when the last push / merge ended with a counter read; else a (small) upper bound: the
    // packed header then carries the exact count and this stays ONE synchronisation
    uint32_t n = counters_stale_ ? (uint32_t)std::min<uint64_t>(groups_bound_, table_->capacity) : groups_known_;
    const int words = 3 + K + W;
    BufPtr packed = dev_alloc(ctx_, (size_t)(n + 1) * words * 8);
    SQ_CUDA(cudaMemsetAsync(packed->p, 0, (size_t)words * 8, ctx_.stream));
    if (n > 256 && !counters_stale_) pack_sorted(K, W, n, (uint64_t*)packed->p);  // ordered on the device
    else launch_table_pack(table_->view(), K, W, (uint64_t*)packed->p, n, ctx_.stream);                              // few groups: the host sorts
    const size_t host_words = (size_t)(n + 1) * words;
    if (pinned_words_ < host_words) {  // pinned staging, kept across runs: the D2H runs at PCIe speed
      if (pinned_) cudaFreeHost(pinned_);
      pinned_ = nullptr;
      pinned_words_ = 0;
      SQ_CUDA(cudaHostAlloc((void**)&pinned_, host_words * 8, cudaHostAllocDefault));
      pinned_words_ = host_words;
    }
    uint64_t* host = pinned_;
    SQ_CUDA(cudaMemcpyAsync(host, packed->p, host_words * 8, cudaMemcpyDeviceToHost, ctx_.stream));
    ctx_.sync();
    if (counters_stale_) {
      if (host[0] > n) fail(SQLRS_ERR_INTERNAL, "group count exceeds its bound under finalisation");
      n = (uint32_t)host[0];
      groups_known_ = n;
      groups_bound_ = n;
      counters_stale_ = false;
    } else if (host[0] != n) {
      fail(SQLRS_ERR_INTERNAL, "group count changed under finalisation");
    }
    g->n = n;
    g->hash.resize(n);
    g->min_row.resize(n);
    g->keys.resize((size_t)n * std::max(K, 1));
    g->knull.resize(n);
    g->acc.resize((size_t)n * std::max(W, 1));
    for (uint32_t i = 0; i < n; i++) {
      const uint64_t* row = host + (size_t)(1 + i) * words;
      g->hash[i] = row[0];
      g->min_row[i] = row[1];
      g->knull[i] = (uint32_t)row[2];
      for (int k = 0; k < K; k++) g->keys[(size_t)k * n + i] = row[3 + k];
      for (int w = 0; w < W; w++) g->acc[(size_t)w * n + i] = row[3 + K + w];
    }
  }
  fields->clear();
  for (int k = 0; k < K; k++) fields->push_back(Field{k < (int)group_names_.size() ? group_names_[k] : "", c.key_dtypes[k], true});
  for (size_t j = 0; j < aggs_.size(); j++) fields->push_back(Field{aggs_[j].name, c.aggs[j].out_dtype, true});
}

static double sortable_to_f64(int64_t s) {
  int64_t b = s ^ ((s >> 63) & 0x7fffffffffffffffLL);
  double d;
  std::memcpy(&d, &b, 8);
  return d;
}

void AggOp::finish_host(ArrowArray* out, ArrowSchema* out_schema) {
  Trace tr("agg.finish_host", ctx_.stream);
  if (hint_sized_) settle();  // a deferred push_join: the group count decides the path below
  // many groups: finalise on the device and copy whole columns (the row-at-a-time host loop below cost 2.7 ms for
  // Q3' SF10's 113 k groups); few groups: one packed D2H and a trivial host loop beat the extra launches
  bool any_utf8 = false;
  if (!cache_.empty()) {
    const Compiled& cc = *cache_.begin()->second;
    for (int dt : cc.key_dtypes) any_utf8 |= dt == SQLRS_DT_UTF8;
    for (const AggPlan& ap : cc.aggs) any_utf8 |= ap.out_dtype == SQLRS_DT_UTF8;
  }
  if (any_utf8 && table_ && counters_stale_) settle();
  if (distinct_ || any_utf8 || (seen_batch_ && table_ && !counters_stale_ && groups_known_ > 1024)) {
    DBatch b = finish_device();
    export_batch_host(ctx_, b, out, out_schema);
    return;
  }
  std::vector<Field> fields;
  HostGroups g;
  {
    Trace t1("  finish.build_output", ctx_.stream);
    build_output(&fields, &g);
  }
  Trace t2("  finish.convert+export", ctx_.stream);
  const Compiled& c = *cache_.begin()->second;
  const int K = (int)c.key_dtypes.size(), W = (int)c.words.size();
  const uint32_t n = g.n;
  // first-appearance order (hash_agg.rs:98,134) = ascending first global row id
  std::vector<uint32_t> order(n);
  for (uint32_t i = 0; i < n; i++) order[i] = i;
  if (!std::is_sorted(g.min_row.begin(), g.min_row.end()))  // large tables arrive ordered from the device (table_pack_sorted)
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return g.min_row[a] < g.min_row[b]; });
  const bool synth_row = simple_ && n == 0;  // SimpleAgg over batches without a surviving row: one row of initial values
  const int64_t rows = synth_row ? 1 : n;
  std::vector<HostCol> cols(fields.size());
  for (int k = 0; k < K; k++) {
    HostCol& col = cols[k];
    col.dtype = c.key_dtypes[k];
    if (col.dtype == SQLRS_DT_NULL)
      fail(SQLRS_ERR_ARROW, "NotYetImplemented: not support Null as group by key");  // types/mod.rs:241-245
    bool any_null = false;
    for (uint32_t i = 0; i < n; i++) any_null |= (g.knull[order[i]] >> k) & 1u;
    if (any_null) col.valid.assign(n, 1);
    (col.dtype == SQLRS_DT_FLOAT64) ? col.f.reserve(n) : col.i.reserve(n);
    for (uint32_t i = 0; i < n; i++) {
      const uint32_t e = order[i];
      const uint64_t bits = g.keys[(size_t)k * n + e];
      const bool is_null = (g.knull[e] >> k) & 1u;
      if (is_null) col.valid[i] = 0;
      if (col.dtype == SQLRS_DT_FLOAT64) {
        double d;
        std::memcpy(&d, &bits, 8);
        col.f.push_back(is_null ? 0.0 : d);
      } else {
        col.i.push_back(is_null ? 0 : (int64_t)bits);
      }
    }
  }
  for (size_t j = 0; j < aggs_.size(); j++) {
    HostCol& col = cols[K + j];
    const AggPlan& p = c.aggs[j];
    col.dtype = p.out_dtype;
    std::vector<uint8_t> valid(rows, 1);
    bool any_null = false;
    (col.dtype == SQLRS_DT_FLOAT64) ? col.f.reserve(rows) : col.i.reserve(rows);
    for (int64_t i = 0; i < rows; i++) {
      uint64_t word = 0, nvalid = 1;
      if (!synth_row) {
        const uint32_t e = order[i];
        word = g.acc[(size_t)p.value_word * n + e];
        if (p.nvalid_word >= 0) nvalid = g.acc[(size_t)p.nvalid_word * n + e];
      } else {
        word = word_identity(c.words[p.value_word].op);
        nvalid = 0;
      }
      if (p.func == SQLRS_AGG_COUNT) {
        uint64_t cnt = word;
        if (c.words[p.value_word].op == W_COUNT_EPOCH) {
          cnt = word & kEpochMask;
          // SimpleAgg updates its single accumulator set with EVERY batch, empty ones included
          // (simple_agg.rs:34-54), so under the overwrite quirk only the last batch counts
          if (simple_ && (word >> kEpochShift) != (uint64_t)batches_seen_) cnt = 0;
        }
        col.i.push_back((int64_t)cnt);
        continue;
      }
      const bool is_null = nvalid == 0;
      if (is_null) {
        valid[i] = 0;
        any_null = true;
      }
      if (col.dtype == SQLRS_DT_FLOAT64) {
        double d;
        if (p.f64_sortable) d = sortable_to_f64((int64_t)word);
        else std::memcpy(&d, &word, 8);
        col.f.push_back(is_null ? 0.0 : d);
      } else {
        col.i.push_back(is_null ? 0 : (int64_t)word);
      }
    }
    if (any_null) col.valid = valid;
  }
  (void)W;
  export_host_columns(fields, cols, rows, out, out_schema);
}

// the same result as finish_host, but as a device-resident batch (columns in HBM): what an operator above the
// aggregate in the same plan consumes (Order / Project / Limit), and the fast path to the host for many groups —
// one kernel turns the ordered packed rows into typed columns + validity words, the host never touches a row
DBatch AggOp::finish_device(DCol* first_row) {
  Trace tr("agg.finish_device", ctx_.stream);
  KernelEvent ev(opt_.flags, ctx_.stream, "aggregate finalise (pack + k_finalize_groups)");
  if (!seen_batch_) fail(SQLRS_ERR_INTERNAL, "called `Option::unwrap()` on a `None` value (no input batch)");
  ctx_.activate();
  if (distinct_) {
    if (first_row) fail(SQLRS_ERR_INTERNAL, "finish_device: DISTINCT aggregates are finalised in order");
    return finish_distinct();
  }
  const Compiled& c = *cache_.begin()->second;
  const int K = (int)c.key_dtypes.size(), W = (int)c.words.size();
  const int words = 3 + K + W;
  if (table_ && counters_stale_) {
    uint32_t hc[4];
    read_counters(hc);
  }
  const uint32_t n = table_ ? groups_known_ : 0;
  const bool synth_row = simple_ && n == 0;  // SimpleAgg over batches without a surviving row: one row of initial values
  const int64_t rows = synth_row ? 1 : n;
  BufPtr packed = dev_alloc(ctx_, (size_t)(rows + 1) * words * 8);
  if (synth_row) {
    std::vector<uint64_t> host((size_t)2 * words, 0);
    for (int w = 0; w < W; w++) host[(size_t)words + 3 + K + w] = word_identity(c.words[w].op);
    SQ_CUDA(cudaMemcpyAsync(packed->p, host.data(), host.size() * 8, cudaMemcpyHostToDevice, ctx_.stream));
    SQ_CUDA(cudaStreamSynchronize(ctx_.stream));  // `host` is a stack vector
  } else if (n > 0) {
    SQ_CUDA(cudaMemsetAsync(packed->p, 0, (size_t)words * 8, ctx_.stream));
    if (first_row) {  // any order: straight from the slot list / a scan of the table
      if (slot_list_complete_) launch_table_pack_list(table_->view(), K, W, (const uint32_t*)table_->new_slots->p, n, (uint64_t*)packed->p, ctx_.stream);
      else launch_table_pack(table_->view(), K, W, (uint64_t*)packed->p, n, ctx_.stream);
    } else {
      // first-appearance order (hash_agg.rs:98,134)
      pack_sorted(K, W, n, (uint64_t*)packed->p);
    }
  }
  DBatch out;
  out.n = rows;
  std::vector<FinalizeCol> desc;
  for (int k = 0; k < K; k++) {
    if (c.key_dtypes[k] == SQLRS_DT_NULL)
      fail(SQLRS_ERR_ARROW, "NotYetImplemented: not support Null as group by key");  // types/mod.rs:241-245
    out.fields.push_back(Field{k < (int)group_names_.size() ? group_names_[k] : "", c.key_dtypes[k], true});
    DCol col = make_col(ctx_, c.key_dtypes[k], rows, c.key_decl_null[k]);
    FinalizeCol d{};
    d.data = col_data(col);
    d.valid = col_valid(col);
    d.dtype = col.dtype;
    d.word = 3 + k;
    d.null_bit = k;
    d.nvalid_word = -1;
    desc.push_back(d);
    out.cols.push_back(col);
  }
  for (size_t j = 0; j < aggs_.size(); j++) {
    const AggPlan& p = c.aggs[j];
    out.fields.push_back(Field{aggs_[j].name, p.out_dtype, true});
    const bool is_count = p.func == SQLRS_AGG_COUNT;
    const bool nullable = !is_count && (p.nvalid_word >= 0 || synth_row);
    DCol col = make_col(ctx_, p.out_dtype, rows, nullable);
    FinalizeCol d{};
    d.data = col_data(col);
    d.valid = col_valid(col);
    d.dtype = col.dtype;
    d.word = 3 + K + p.value_word;
    d.null_bit = -1;
    d.nvalid_word = (!is_count && p.nvalid_word >= 0) ? 3 + K + p.nvalid_word : -1;
    if (synth_row && !is_count) d.nvalid_word = 0;  // word 0 (hash) of the synthetic row is 0: SUM/MIN/MAX of nothing is NULL
    d.f64_sortable = p.f64_sortable ? 1 : 0;
    d.utf8_packed = p.utf8_packed ? 1 : 0;
    if (is_count && c.words[p.value_word].op == W_COUNT_EPOCH) {
      d.count_epoch = 1;
      // SimpleAgg updates its single accumulator set with EVERY batch, empty ones included (simple_agg.rs:34-54)
      d.simple_epoch = simple_ ? (uint64_t)batches_seen_ : 0;
    }
    desc.push_back(d);
    out.cols.push_back(col);
  }
  if (first_row) {  // the groups' first-appearance ordinals (packed word 1) as one more Int64 column
    *first_row = make_col(ctx_, SQLRS_DT_INT64, rows, false);
    FinalizeCol d{};
    d.data = col_data(*first_row);
    d.valid = nullptr;
    d.dtype = SQLRS_DT_INT64;
    d.word = 1;
    d.null_bit = -1;
    d.nvalid_word = -1;
    desc.push_back(d);
  }
  if (rows > 0 && !desc.empty()) {
    launch_finalize_groups((const uint64_t*)packed->p, words, rows, (int)desc.size(), desc.data(), ctx_.stream);
    ctx_.defer([packed]() {});
  }
  return out;
}

// DISTINCT: zip of the plain aggregates with one second-level aggregate per DISTINCT aggregate (see struct Distinct)
DBatch AggOp::finish_distinct() {
  Distinct& d = *distinct_;
  const size_t K = group_by_.size();
  Options sub = opt_;
  sub.stream = ctx_.stream;
  sub.device_id = ctx_.device;
  sub.count_mode = SQLRS_COUNT_SQL_ACCUMULATE;  // the second level counts set elements; quirk K1 is about input batches
  sub.match_mode = SQLRS_MATCH_HASH_AND_KEY;
  DBatch plain;  // [keys..., plain aggregates...]
  if (d.plain) plain = d.plain->finish_device();
  else plain.n = 1;
  DBatch out;
  out.n = plain.n;
  for (size_t k = 0; k < K; k++) {
    out.fields.push_back(plain.fields[k]);
    out.cols.push_back(plain.cols[k]);
  }
  std::vector<DCol> agg_cols(aggs_.size());
  std::vector<Field> agg_fields(aggs_.size());
  std::vector<bool> filled(aggs_.size(), false);
  for (auto& it : d.items) {
    const AggSpec& a = aggs_[it.agg_index];
    DBatch elems = it.dedup->finish_device();  // [keys..., x], one row per distinct (keys, x), first-appearance order
    std::vector<ExprCopy> keys;
    for (size_t k = 0; k < K; k++) {
      ExprNodeCopy n;
      n.op = SQLRS_OP_INPUT_REF;
      n.index = (int)k;
      n.dtype = elems.cols[k].dtype;
      keys.push_back(ExprCopy{n});
    }
    AggSpec second;
    second.func = a.func;
    second.distinct = 0;
    second.name = a.name;
    ExprNodeCopy arg;
    if (a.func == SQLRS_AGG_COUNT) {  // the set's size: a NULL element counts (count.rs:44-57)
      arg.op = SQLRS_OP_CONSTANT;
      arg.dtype = SQLRS_DT_INT32;
      arg.imm_bits = 1;
      second.return_dtype = SQLRS_DT_INT64;
    } else {
      arg.op = SQLRS_OP_INPUT_REF;
      arg.index = (int)K;
      arg.dtype = elems.cols[K].dtype;
      second.return_dtype = a.return_dtype;
    }
    second.arg = ExprCopy{arg};
    std::vector<std::string> names(group_names_.begin(), group_names_.end());
    names.resize(K);
    AggOp op2(std::vector<AggSpec>{second}, keys, names, simple_, ExprCopy(), sub);
    op2.push(elems);
    DBatch r = op2.finish_device();  // [keys..., aggregate]
    if (r.n != plain.n)
      fail(SQLRS_ERR_UNSUPPORTED, "DISTINCT aggregate: the hash-only group identity merged groups that differ by key (quirk K2)");
    agg_cols[it.agg_index] = r.cols[K];
    agg_fields[it.agg_index] = Field{a.name, r.cols[K].dtype, true};
    filled[it.agg_index] = true;
  }
  size_t p = K;
  for (size_t j = 0; j < aggs_.size(); j++) {
    if (!filled[j]) {
      agg_cols[j] = plain.cols[p];
      agg_fields[j] = plain.fields[p];
      p++;
    }
    out.fields.push_back(agg_fields[j]);
    out.cols.push_back(agg_cols[j]);
  }
  return out;
}

// ------------------------------------------------------------------ partial / final (multi-GPU)
void AggOp::set_row_base(int64_t first_global_row) {
  rows_seen_ = first_global_row;
  if (distinct_) {
    if (distinct_->plain) distinct_->plain->set_row_base(first_global_row);
    for (auto& it : distinct_->items) it.dedup->set_row_base(first_global_row);
  }
}

// the tables of a DISTINCT composition: [plain aggregates (if any), one dedup table per DISTINCT aggregate in declared order]
int AggOp::partial_tables() const {
  if (!distinct_) return 1;
  return (distinct_->plain ? 1 : 0) + (int)distinct_->items.size();
}
AggOp& AggOp::partial_table(int index) {
  if (index < 0 || index >= partial_tables()) fail(SQLRS_ERR_INVALID_ARG, "partials table index out of range");
  if (!distinct_) return *this;
  if (distinct_->plain) {
    if (index == 0) return *distinct_->plain;
    index--;
  }
  return *distinct_->items[(size_t)index].dedup;
}

void AggOp::check_partial_supported() const {
  if (distinct_) fail(SQLRS_ERR_INTERNAL, "partial/final DISTINCT aggregates are addressed table by table (partial_table)");
  if (!cache_.empty())
    for (const AggPlan& ap : cache_.begin()->second->aggs)
      if (ap.utf8_packed) fail(SQLRS_ERR_UNSUPPORTED, "partial/final MIN / MAX over Utf8 (the packed ranks are local to one process's string pool state)");
  if (opt_.count_mode == SQLRS_COUNT_REFERENCE_OVERWRITE)
    for (const AggSpec& a : aggs_)
      if (a.func == SQLRS_AGG_COUNT)
        fail(SQLRS_ERR_UNSUPPORTED, "partial/final COUNT needs SQLRS_COUNT_SQL_ACCUMULATE (the overwrite quirk K1 is defined on one batch stream)");
}

int AggOp::partial_row_words() const {
  if (cache_.empty()) fail(SQLRS_ERR_INVALID_ARG, "no batch aggregated yet (accumulator layout unknown)");
  const Compiled& c = *cache_.begin()->second;
  return 3 + (int)c.key_dtypes.size() + (int)c.words.size();
}

// The un-finalised group table as a host batch: [hash i64, min_row i64, knull i32, key bits i64 x K,
// accumulator words i64 x W].  Column 0 is the group identity the exchange radix-partitions on.
void AggOp::export_partials(ArrowArray* out, ArrowSchema* out_schema) {
  check_partial_supported();
  std::vector<Field> fields;
  HostGroups g;
  build_output(&fields, &g);
  const Compiled& c = *cache_.begin()->second;
  const int K = (int)c.key_dtypes.size(), W = (int)c.words.size();
  const uint32_t n = g.n;
  std::vector<Field> pf;
  std::vector<HostCol> cols;
  auto add_i64 = [&](const std::string& name, const uint64_t* src) {
    HostCol col;
    col.dtype = SQLRS_DT_INT64;
    col.i.assign((const int64_t*)src, (const int64_t*)src + n);
    cols.push_back(std::move(col));
    pf.push_back(Field{name, SQLRS_DT_INT64, false});
  };
  add_i64("hash", g.hash.data());
  add_i64("min_row", g.min_row.data());
  {
    HostCol col;
    col.dtype = SQLRS_DT_INT32;
    for (uint32_t i = 0; i < n; i++) col.i.push_back((int64_t)g.knull[i]);
    cols.push_back(std::move(col));
    pf.push_back(Field{"knull", SQLRS_DT_INT32, false});
  }
  for (int k = 0; k < K; k++) add_i64("key" + std::to_string(k), g.keys.data() + (size_t)k * n);
  for (int w = 0; w < W; w++) add_i64("acc" + std::to_string(w), g.acc.data() + (size_t)w * n);
  export_host_columns(pf, cols, n, out, out_schema);
}

// the same, packed row-major, written into caller-provided DEVICE memory on the operator's stream
// (no host round trip): (cap_rows + 1) rows of partial_row_words() u64; row 0 = header {group count}
void AggOp::export_partials_device(uint64_t* dst, int64_t cap_rows) {
  check_partial_supported();
  if (!seen_batch_) fail(SQLRS_ERR_INTERNAL, "called `Option::unwrap()` on a `None` value (no input batch)");
  ctx_.activate();
  const int words = partial_row_words();
  SQ_CUDA(cudaMemsetAsync(dst, 0, (size_t)words * 8, ctx_.stream));
  if (table_ && cap_rows > 0) launch_table_pack(table_->view(), table_->n_keys, table_->n_acc, dst, (uint64_t)cap_rows, ctx_.stream, tier_pending_);
}

int64_t AggOp::export_partials_partitioned(uint64_t* dst, int n_parts, int64_t cap_rows) {
  check_partial_supported();
  if (!seen_batch_) fail(SQLRS_ERR_INTERNAL, "called `Option::unwrap()` on a `None` value (no input batch)");
  if (n_parts < 1) fail(SQLRS_ERR_INVALID_ARG, "n_parts must be >= 1");
  ctx_.activate();
  uint32_t hc[4] = {0, 0, 0, 0};
  if (table_) read_counters(hc);
  const int64_t groups = table_ ? (int64_t)groups_known_ : 0;
  if (cap_rows <= 0 || !dst) return groups;
  const int words = partial_row_words();
  for (int q = 0; q < n_parts; q++) SQ_CUDA(cudaMemsetAsync(dst + (size_t)q * (cap_rows + 1) * words, 0, (size_t)words * 8, ctx_.stream));
  if (table_ && groups > 0) launch_table_pack_partitioned(table_->view(), table_->n_keys, table_->n_acc, dst, n_parts, (uint64_t)cap_rows, ctx_.stream);
  return groups;
}

void AggOp::clear_partials() {
  ctx_.activate();
  tier_pending_ = false;  // (the state the check was about is discarded; a packed export already carries the overflow mark)
  slot_list_complete_ = false;
  if (table_) init_table_contents(*table_);
  level_ = 0;
  groups_known_ = 0;
  groups_bound_ = 0;
  counters_stale_ = false;
}

const int* AggOp::device_word_ops() {
  if (d_ops_) return (const int*)d_ops_->p;
  const Compiled& c = *cache_.begin()->second;
  std::vector<int> ops;
  for (const WordPlan& w : c.words) {
    switch (w.op) {
      case W_ADD_U64: ops.push_back(0); break;
      case W_ADD_F64: ops.push_back(1); break;
      case W_MIN_I64: ops.push_back(2); break;
      case W_MAX_I64: ops.push_back(3); break;
      case W_COUNT_EPOCH: ops.push_back(4); break;  // batch-local tables of push_join; the multi-GPU export refuses this mode
      default: fail(SQLRS_ERR_INTERNAL, "unknown accumulator word");
    }
  }
  d_ops_ = dev_alloc(ctx_, std::max<size_t>(ops.size(), 1) * 4);
  if (!ops.empty()) {
    SQ_CUDA(cudaMemcpyAsync(d_ops_->p, ops.data(), ops.size() * 4, cudaMemcpyHostToDevice, ctx_.stream));
    SQ_CUDA(cudaStreamSynchronize(ctx_.stream));  // `ops` is a stack vector
  }
  return (const int*)d_ops_->p;
}

// folds n_bufs packed partial buffers (device memory, layout of export_partials_device) into the table
void AggOp::merge_partials_device(const uint64_t* src, int n_bufs, int64_t cap_rows, bool sync_after) {
  ctx_.activate();
  slot_list_complete_ = false;
  if (cache_.empty()) fail(SQLRS_ERR_INVALID_ARG, "merge_partials before any batch was aggregated (accumulator layout unknown)");
  seen_batch_ = true;
  if (n_bufs <= 0 || cap_rows <= 0) return;
  const int* ops = device_word_ops();
  const uint64_t extra = (uint64_t)n_bufs * (uint64_t)cap_rows;
  if (!table_) {
    ensure_table((uint32_t)std::min<uint64_t>(2ULL * extra + 1024, 1ULL << 31));
  } else if (2ULL * (groups_bound_ + extra) + 1024 > table_->capacity) {
    uint32_t hc[4];
    read_counters(hc);
    ensure_table((uint32_t)std::min<uint64_t>(2ULL * (groups_known_ + extra) + 1024, 1ULL << 31));
  }
  launch_table_merge_packed(table_->view(), table_->n_keys, table_->n_acc, ops, opt_.match_mode == SQLRS_MATCH_HASH_AND_KEY ? 1 : 0, src, n_bufs,
                            (uint64_t)cap_rows, ctx_.stream);
  groups_bound_ += extra;
  if (sync_after) {
    uint32_t hc[4];
    read_counters(hc);
  } else {
    counters_stale_ = true;  // the next reader of the group count fetches it (build_output: with the packed result)
  }
}

// folds a batch of partial groups (layout of export_partials, device resident columns) into the table
void AggOp::merge_partials(const DBatch& p) {
  ctx_.activate();
  if (cache_.empty()) fail(SQLRS_ERR_INVALID_ARG, "merge_partials before any batch was aggregated (accumulator layout unknown)");
  const int words = partial_row_words();
  if ((int)p.cols.size() != words) fail(SQLRS_ERR_INVALID_ARG, "partials batch has the wrong number of columns");
  for (size_t k = 0; k < p.cols.size(); k++)
    if (p.cols[k].dtype != (k == 2 ? SQLRS_DT_INT32 : SQLRS_DT_INT64) || p.cols[k].valid)
      fail(SQLRS_ERR_INVALID_ARG, "partials batch has the wrong column types");
  const int64_t n = p.n;
  seen_batch_ = true;
  if (n == 0) return;
  // columns -> one packed buffer (strided copies), header = n
  BufPtr packed = dev_alloc_zero(ctx_, (size_t)(n + 1) * words * 8);
  uint64_t header = (uint64_t)n;
  SQ_CUDA(cudaMemcpyAsync(packed->p, &header, 8, cudaMemcpyHostToDevice, ctx_.stream));
  for (int k = 0; k < words; k++) {
    const size_t w = k == 2 ? 4 : 8;
    SQ_CUDA(cudaMemcpy2DAsync((uint64_t*)packed->p + words + k, (size_t)words * 8, p.cols[k].data, w, w, (size_t)n, cudaMemcpyDeviceToDevice,
                              ctx_.stream));
  }
  merge_partials_device((const uint64_t*)packed->p, 1, n, true);  // ends with a synchronising counter read: `header` stays valid
}

}  // namespace sq
