// sqlrs_b200 — the operators that follow the hot path in a v1 plan, on device batches:
//   ProjectOp  ProjectExecutor  reference src/executor/project.rs:6-29
//   OrderOp    OrderExecutor    reference src/executor/order.rs:8-67
//   LimitOp    LimitExecutor    reference src/executor/limit.rs:6-80
// The planner stacks them Agg -> Order -> Project -> Limit (src/planner/select.rs:34-45); keeping them on the device
// means only the final rows (Q3': 10) cross PCIe instead of the whole aggregate output.
#pragma once
#include "ops.hpp"

namespace sq {

class ProjectOp {
 public:
  // keep_field[k]: names[k] was NULL at the ABI — a bare InputRef keeps the input field (evaluator.rs:31)
  ProjectOp(std::vector<ExprCopy> exprs, std::vector<std::string> names, std::vector<bool> keep_field, const Options& opt);
  DBatch execute(const DBatch& in);
  DBatch execute(Ctx& ctx, const DBatch& in);  // on another operator's / the plan's context
  Ctx& ctx() { return ctx_; }

 private:
  Ctx ctx_;
  std::vector<ExprCopy> exprs_;
  std::vector<std::string> names_;
  std::vector<bool> keep_field_;
  std::unique_ptr<EvalProgram> prog_;  // the non-trivial expressions, one fused kernel
  std::vector<int> slot_;              // per expression: index into prog_'s outputs, or -1 = bare InputRef (column passed through)
};

class OrderOp {
 public:
  OrderOp(std::vector<ExprCopy> order_by, std::vector<bool> asc, const Options& opt);
  void push(const DBatch& b) { batches_.push_back(b); }
  // plan-level fusion with a Limit above (top-k): only the first `rows` sorted rows are gathered; < 0 = all
  void set_row_limit(int64_t rows) { row_limit_ = rows; }
  // Rows that tie on every sort key come out in input order (stable).  An aggregate below may hand its groups over in
  // ANY order together with this column of their first-appearance ordinals (unique, Int64): ties then resolve by it,
  // which is the order the sorted hand-over would have produced — without sorting the groups first.
  void set_tiebreak(DCol first_row) { tiebreak_ = std::move(first_row); has_tiebreak_ = true; }
  // can finish() use the top-k selection (no full sort) for this row limit?  (what the plan asks before it skips the
  // aggregate's ordered finalisation)
  bool topk_applies(int64_t row_limit) const;
  DBatch finish();
  DBatch finish(Ctx& ctx);
  void reset() {
    batches_.clear();
    has_tiebreak_ = false;
    tiebreak_ = DCol();
  }
  Ctx& ctx() { return ctx_; }

 private:
  Ctx ctx_;
  std::vector<ExprCopy> order_by_;
  std::vector<bool> asc_;
  std::vector<DBatch> batches_;
  std::unique_ptr<EvalProgram> prog_;
  std::vector<int> slot_;
  int64_t row_limit_ = -1;
  int flags_ = 0;
  DCol tiebreak_;
  bool has_tiebreak_ = false;
};

class LimitOp {
 public:
  LimitOp(int64_t limit, int64_t offset, const Options& opt);  // -1 = None
  bool push(const DBatch& in, DBatch* out);                     // true: *out is yielded
  bool push(Ctx& ctx, const DBatch& in, DBatch* out);
  bool done() const { return done_; }
  void reset() {
    returned_count_ = 0;
    done_ = limit_ == 0;
  }
  // rows of the child's stream that can reach the output when the child yields ONE batch (Order below): offset + limit
  int64_t rows_needed() const { return limit_ < 0 ? -1 : (offset_ < 0 ? 0 : offset_) + limit_; }
  Ctx& ctx() { return ctx_; }

 private:
  Ctx ctx_;
  int64_t limit_, offset_, returned_count_ = 0;
  bool done_ = false;
};

// CrossJoinExecutor, reference src/executor/join/cross_join.rs:8-57: the left side is drained and concatenated; every
// right batch yields ONE output batch per left row (that row repeated to the right batch's length + the right columns,
// which are shared, not copied).  Built from the take / fill primitives only — a cross join is not on the hot path.
class CrossJoinOp {
 public:
  CrossJoinOp(std::vector<Field> out_fields, const Options& opt);
  void build_push(const DBatch& b);
  std::vector<DBatch> probe(const DBatch& right);
  std::vector<DBatch> probe(Ctx& ctx, const DBatch& right);
  Ctx& ctx() { return ctx_; }

 private:
  Ctx ctx_;
  std::vector<Field> out_fields_;
  std::vector<DBatch> left_batches_;
  bool sealed_ = false, has_left_ = false;
  DBatch left_single_;
};

// rows [start, start + len) of a batch as a new batch (RecordBatch::slice materialised)
DBatch slice_batch(Ctx& ctx, const DBatch& in, int64_t start, int64_t len);
DBatch concat_batches(Ctx& ctx, const std::vector<DBatch>& batches);

}  // namespace sq
