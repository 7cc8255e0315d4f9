// sqlrs_b200 — a physical sub-plan on the GPU: what ExecutorBuilder::build(plan) wires together in
// the reference (src/executor/mod.rs:45-47, visit_* :87-200).  Tables stay resident in HBM, the
// intermediate batches never leave the device, and Filter directly below an aggregate is fused
// into the aggregate's row program (one pass over the scan, nothing materialised).
#pragma once
#include <deque>
#include <map>

#include "join.hpp"
#include "ops.hpp"
#include "tail.hpp"

namespace sq {

class Plan {
 public:
  Plan(const sqlrs_plan_node* nodes, int32_t n_nodes, int32_t root, const Options& opt);
  ~Plan() { reset(); }
  void push_table(int slot, DBatch batch);
  void execute();
  bool next(ArrowArray* out, ArrowSchema* out_schema);
  void reset();
  void clear_table(int slot);
  void execute_partial(int64_t row_base);
  void export_partials(ArrowArray* out, ArrowSchema* out_schema);
  void clear_partials();
  void merge_partials(const DBatch& partials);
  void finish_partial();
  int partial_row_words() const;
  int partials_tables() const;  // > 1 with DISTINCT aggregates; the partial-state calls address the selected table
  void select_partials_table(int index);
  void export_partials_device(uint64_t* dst, int64_t cap_rows);
  int64_t export_partials_partitioned(uint64_t* dst, int n_parts, int64_t cap_rows);
  bool result_shape(int64_t* n_rows, int32_t* n_cols);
  void next_to_device(void* const* columns, int32_t n_columns);
  void push_table_batched(int slot, const DBatch& whole, int64_t batch_rows);
  void merge_partials_device(const uint64_t* src, int n_bufs, int64_t cap_rows);
  const char* describe() const { return description_.c_str(); }
  double scan_kernel_ms();
  int64_t scan_kernel_launches();
  Ctx& ctx() { return ctx_; }

 private:
  struct Node {
    int kind = 0, child0 = -1, child1 = -1, table_slot = 0, join_type = 0;
    ExprCopy predicate;
    std::vector<AggSpec> aggs;
    std::vector<ExprCopy> group_by, left_keys, right_keys;
    std::vector<std::string> group_names;
    std::vector<Field> join_fields;
    // operators kept across execute() calls so that repeated runs reuse compiled kernels
    std::unique_ptr<EvalProgram> filter_prog;
    std::unique_ptr<AggOp> agg_op;
    std::unique_ptr<ProjectOp> project_op;
    std::unique_ptr<OrderOp> order_op;
    std::unique_ptr<JoinChainOp> chain_op;  // HASH_JOIN whose build side is another join: fused probe -> build (join.hpp)
    // PROJECT: select list + field names; ORDER: sort expressions + directions; LIMIT: the bound constants (-1 = None)
    std::vector<ExprCopy> exprs;
    std::vector<std::string> expr_names;
    std::vector<bool> asc, keep_field;
    int64_t limit = -1, offset = -1;
    int64_t row_limit_hint = -1;  // ORDER: set by a Limit above for the current run (top-k: gather only these rows)
  };
  struct Result {
    bool on_host = false;
    ArrowArray arr{};
    ArrowSchema sch{};
    DBatch dev;
  };
  using Needed = std::vector<bool>;  // per output column of a node: does anything above read it?  empty = all
  std::vector<DBatch> run(int idx, const Needed& needed);
  void run_agg_to_host(int idx, Result* res);
  AggOp& run_agg(int idx);  // everything of an aggregate node up to (excluding) finalisation
  bool feed_fused_join(AggOp& op, int child, const ExprCopy& agg_fused_pred, const Needed& need);
  // join `jidx`'s build side is the inner join `left`: let that join's probe build j's table (JoinChainOp)
  bool try_chain(int jidx, int left, const Needed& left_need, JoinOp& j);
  // runs `body` until the hint-sized (unsynchronised) parts of the run validate; see JoinChainOp / AggOp::push_join
  void run_validated(const std::function<void()>& body);
  int width_of(int idx);              // number of output columns of a node (needs its scans pushed)
  Needed agg_child_needs(const Node& agg, const ExprCopy& fused_pred, int child_width) const;
  bool fusion() const { return !(opt_.flags & SQLRS_FLAG_NO_FUSION); }

  Ctx ctx_;
  Options opt_;
  std::vector<Node> nodes_;
  int root_ = 0;
  std::map<int, std::vector<DBatch>> tables_;
  std::deque<Result> results_;
  std::unique_ptr<AggOp> partial_op_;  // root aggregate of execute_partial(), kept across runs
  bool partial_active_ = false;        // between execute_partial and finish_partial
  int partial_sel_ = 0;                // table of partial_op_ the partial-state calls address
  int64_t partial_row_base_ = 0;
  bool partial_defer_ok_ = true;       // execute_partial may leave the first scan launch unchecked (settle_partial)
  bool scan_from_partial_ = false;     // scan_kernel_ms() reports partial_op_'s timers
  void settle_partial();
  std::string description_;
  std::vector<JoinChainOp*> pending_chains_;  // runs sized by hints, to be validated once the stream has been synchronised
  double scan_kernel_ms_ = 0;
  int64_t scan_kernel_launches_ = 0;
};

}  // namespace sq
