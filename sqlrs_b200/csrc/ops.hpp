// sqlrs_b200 — the operators behind the C ABI, working on device batches.
//   EvalProgram  BoundExpr::eval_column            reference src/executor/evaluator.rs:13-28
//   FilterOp     FilterExecutor                    src/executor/filter.rs:7-26
//   AggOp        SimpleAggExecutor/HashAggExecutor src/executor/aggregate/{simple_agg.rs:10-65,hash_agg.rs:15-150}
//   JoinOp       HashJoinExecutor                  src/executor/join/hash_join.rs:16-323
#pragma once
#include <map>

#include "codegen.hpp"
#include "device.hpp"
#include "jit.hpp"

namespace sq {

std::vector<ColInfo> col_infos(const DBatch& b);
std::string gen_input_decls(const std::vector<ColInfo>& cols);

// ---- generic "evaluate expressions into columns" program -----------------------------------
enum OutKind {
  OUT_VALUE = 0,  // typed column with validity
  OUT_KEEP = 1,   // Boolean "valid && true" mask, bit-packed, never NULL (the Filter keep-mask)
  OUT_HASH = 2,   // u64 create_hashes() over ALL expressions of the request marked is_key
  OUT_RAWBITS = 3,// u64 raw key bits of one expression (NULL -> 0)
  OUT_NULLMASK = 4,// u32 null mask over the is_key expressions
  OUT_MIXHASH = 5 // u64 cheap multiply-xorshift hash over the is_key expressions' raw bits + types + null flags: a PLACEMENT
                  // hash for tables that also compare the key tuples (SQLRS_MATCH_HASH_AND_KEY); ~4x fewer instructions
                  // than the reference's folded-multiply hash_one
};
struct EvalRequest {
  std::vector<ExprCopy> exprs;
  std::vector<bool> is_key;      // per expr: participates in OUT_HASH / OUT_NULLMASK
  struct Out {
    int kind;
    int expr;  // index into exprs (OUT_VALUE / OUT_KEEP / OUT_RAWBITS)
  };
  std::vector<Out> outs;
};
struct EvalResult {
  std::vector<DCol> cols;           // one per requested output (u64/u32 outputs typed INT64/INT32)
  std::vector<int> expr_dtypes;     // static dtype of every expression
};
class EvalProgram {
 public:
  explicit EvalProgram(EvalRequest req) : req_(std::move(req)) {}
  EvalResult run(Ctx& ctx, const DBatch& batch, const char* what);
  std::string source_for(const std::vector<ColInfo>& cols, std::vector<int>* out_dtypes, std::vector<bool>* out_nullable,
                         std::vector<int>* expr_dtypes);

 private:
  struct Compiled {
    JitKernel* kernel = nullptr;
    std::vector<int> out_dtypes, expr_dtypes;
    std::vector<bool> out_nullable;
  };
  EvalRequest req_;
  std::map<std::string, Compiled> cache_;
};

void check_error_flag(Ctx& ctx, const BufPtr& err, const char* what);  // syncs; throws ERR_ARROW "Divide by zero error"

// take / filter helpers on whole batches
DCol gather_col_u32(Ctx& ctx, const DCol& src, const uint32_t* idx, int64_t m);
DCol gather_col_i64(Ctx& ctx, const DCol& src, const int64_t* idx, int64_t m, bool idx_may_be_null);
// keep-bitmap -> (device u32 index list, count); synchronises to learn the count
BufPtr compact_indices(Ctx& ctx, const uint32_t* keep, int64_t n, int64_t* out_count);
DCol concat_cols(Ctx& ctx, const std::vector<DCol>& parts, int dtype);

// ---- Filter -------------------------------------------------------------------------------------
class FilterOp {
 public:
  FilterOp(const ExprCopy& predicate, const Options& opt);
  DBatch execute(const DBatch& in);
  Ctx& ctx() { return ctx_; }

 private:
  Ctx ctx_;
  EvalProgram prog_;
};
// `needed` (optional): gather only these columns, the others become Null-typed placeholders
DBatch filter_batch(Ctx& ctx, EvalProgram& prog, const DBatch& in, const std::vector<bool>* needed = nullptr);

// ---- aggregates ------------------------------------------------------------------------------------
struct AggSpec {
  int func = 0, distinct = 0, return_dtype = 0;
  ExprCopy arg;
  std::string name;
};
std::vector<AggSpec> copy_aggs(const sqlrs_agg_desc* aggs, int32_t n);

class AggOp {
 public:
  // group_by empty + simple = SimpleAggExecutor; `predicate` non-empty = fused Filter below the aggregate
  AggOp(std::vector<AggSpec> aggs, std::vector<ExprCopy> group_by, std::vector<std::string> group_names, bool simple,
        ExprCopy fused_predicate, const Options& opt);
  ~AggOp();
  void push(const DBatch& batch);
  // result as a device batch.  first_row != nullptr: the groups in ANY order (no sort by first appearance) plus the
  // column of their first-appearance ordinals — for a consumer that orders by something else (Order above the aggregate)
  DBatch finish_device(DCol* first_row = nullptr);
  void finish_host(ArrowArray* out, ArrowSchema* out_schema);  // result straight to host Arrow
  Ctx& ctx() { return ctx_; }
  std::string describe() const;
  std::string debug_source(const std::vector<ColInfo>& cols);  // generated CUDA for batches of that schema (no GPU needed)
  std::string debug_join_source(const std::vector<ColInfo>& build_cols, const std::vector<ColInfo>& probe_cols,
                                const std::vector<ExprCopy>& right_keys, const ExprCopy& probe_pred, const ExprCopy& join_filter);
  // partial/final split for multi-GPU execution: the raw group table as a batch
  // [hash, min_row, knull, key bits..., accumulator words...] and its merge into another operator
  void export_partials(ArrowArray* out, ArrowSchema* out_schema);
  void clear_partials();
  void merge_partials(const DBatch& partials);
  int partial_row_words() const;  // u64 words per packed partial row: 3 + keys + accumulator words
  void export_partials_device(uint64_t* dst, int64_t cap_rows);
  int64_t export_partials_partitioned(uint64_t* dst, int n_parts, int64_t cap_rows);  // returns the group count (one sync)
  void merge_partials_device(const uint64_t* src, int n_bufs, int64_t cap_rows, bool sync_after = false);
  void reset();  // forget all groups, keep compiled kernels and buffers
  // fused probe -> aggregate over an INNER hash join whose build side is sealed in `join` (csrc/jit/joinagg.cuh):
  // `probe` is a batch of the join's right child, `probe_pred` a Filter fused below the join on that side
  // `defer`: the caller validates this run afterwards (Plan::run_validated), so when a previous run of this operator left
  // a group-count hint the table is sized from it and NOTHING synchronises here; the counters are read by the next
  // consumer (finish / settle), and a table that turned out too small raises RetrySizingError there
  void push_join(const DBatch& probe, class JoinOp& join, const ExprCopy& probe_pred, bool defer = false);
  void settle();  // reads the counters of a deferred push_join now (one synchronisation)
  void settle_hint_sized() {  // only if the table was sized from a hint whose validity nobody has checked yet
    if (hint_sized_) settle();
  }
  void set_row_base(int64_t first_global_row);
  // partial/final split: let the first batch's sq_agg_small launch go unchecked (no host synchronisation between the scan kernel
  // and the exchange that follows).  Whether its per-CTA slots overflowed is learnt at the next counter read (RetrySizingError:
  // the caller re-runs without deferral); export_partials_device does not read counters: it marks the packed buffer's
  // header (count > any capacity) so that every rank of the exchange sees it.
  void set_defer_tier_check(bool on) { defer_tier_check_ = on && !defer_disabled_; }
  bool tier_check_pending() const { return tier_pending_; }

  bool has_distinct() const { return distinct_ != nullptr; }
  // DISTINCT aggregates keep one group table per DISTINCT aggregate (the set elements) next to the table of the plain
  // aggregates: the partial/final calls above address ONE table, this is how a caller walks all of them
  int partial_tables() const;
  AggOp& partial_table(int index);

 private:
  struct Compiled;
  struct Table;
  struct Distinct;  // DISTINCT aggregates: this operator becomes a composition of DISTINCT-free operators (ops_agg.cpp)
  Compiled& compiled_for(const DBatch& batch);
  struct JoinGen {  // what the fused probe->aggregate program needs to know about the join below
    std::vector<ColInfo> build_cols;
    std::vector<ExprCopy> right_keys;
    ExprCopy probe_pred, join_filter;
    bool jmatch = false;
  };
  std::string generate(const std::vector<ColInfo>& cols, Compiled& comp, const JoinGen* jg = nullptr);
  std::unique_ptr<Table> new_table(uint32_t capacity);
  void init_table_contents(Table& t);
  void ensure_partial_scratch(size_t entries, int K, size_t W);
  void read_counters(uint32_t* out4);
  void check_partial_supported() const;
  const int* device_word_ops();
  void ensure_table(uint32_t min_capacity);
  void grow_table(uint32_t min_capacity);
  void build_output(std::vector<Field>* fields, struct HostGroups* groups);
  void pack_sorted(int K, int W, uint32_t n, uint64_t* dst);
  DBatch finish_distinct();

  Ctx ctx_;
  Options opt_;
  std::vector<AggSpec> aggs_;
  std::vector<ExprCopy> group_by_;
  std::vector<std::string> group_names_;
  bool simple_;
  ExprCopy predicate_;
  std::unique_ptr<Distinct> distinct_;
  std::map<std::string, std::unique_ptr<Compiled>> cache_;
  struct JoinKernels {
    JitKernel* generic = nullptr;
    JitKernel* tma = nullptr;      // sq_joinagg_tma_kernel, when the probe program's columns can be staged by TMA
    std::vector<int> tile_cols;    // the staged probe columns
  };
  std::map<std::string, JoinKernels> join_kernels_;  // fused probe->aggregate kernels by (probe, build) schema signature
  std::unique_ptr<Table> table_;
  int64_t rows_seen_ = 0, batches_seen_ = 0;
  bool seen_batch_ = false;
  int level_ = 0;  // sticky per operator: 0 sq_agg_small, 1 sq_agg_medium, 2 sq_agg_global
  uint32_t groups_known_ = 0;   // exact group count at the last counter read
  uint64_t groups_bound_ = 0;   // host-side upper bound since then
  bool counters_stale_ = false; // device work since the last counter read may have added groups
  uint32_t groups_hint_ = 0;    // group count at the end of the previous run (0 = none)
  bool hint_sized_ = false;     // the current table was sized from groups_hint_ and its overflow flag has not been read yet
  bool slot_list_pending_ = false;  // a deferred push_join: whether new_slots lists every group is known at the counter read
  bool slot_list_complete_ = false;  // table_->new_slots[0 .. groups_known_) lists every occupied slot (table adopted from one
                                     // fused probe+aggregate launch, untouched since): finalisation need not scan the capacity
  uint64_t* pinned_ = nullptr;  // pinned host staging of the packed result
  size_t pinned_words_ = 0;
  size_t part_entries_ = 0;     // CTA-partial scratch of sq_agg_small
  BufPtr p_state_, p_hash_, p_min_, p_keys_, p_knull_, p_acc_, d_ops_;
  std::vector<int> key_dtypes_;
  std::string last_path_;
  double scan_kernel_ms_ = 0;      // SQLRS_FLAG_TIMING: device time of the scan kernels (CUDA events on ctx_.stream)
  int64_t scan_kernel_launches_ = 0;
  bool defer_tier_check_ = false, defer_disabled_ = false, tier_pending_ = false;
  cudaEvent_t pend_e0_ = nullptr, pend_e1_ = nullptr;  // the deferred launch's timer, resolved on demand
  void resolve_pending_timer();

 public:
  double scan_kernel_ms() {
    resolve_pending_timer();
    return scan_kernel_ms_;
  }
  int64_t scan_kernel_launches() {
    resolve_pending_timer();
    return scan_kernel_launches_;
  }
};

}  // namespace sq
