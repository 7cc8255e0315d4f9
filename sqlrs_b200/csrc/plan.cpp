// sqlrs_b200 — physical sub-plan executor (see plan.hpp).
#include "plan.hpp"

namespace sq {

Plan::Plan(const sqlrs_plan_node* nodes, int32_t n_nodes, int32_t root, const Options& opt) : ctx_(opt), opt_(opt), root_(root) {
  opt_.stream = ctx_.stream;  // every operator of the plan works on the plan's stream
  opt_.device_id = ctx_.device;
  for (int32_t k = 0; k < n_nodes; k++) {
    const sqlrs_plan_node& s = nodes[k];
    Node n;
    n.kind = s.kind;
    n.child0 = s.child0;
    n.child1 = s.child1;
    n.table_slot = s.table_slot;
    n.join_type = s.join_type;
    n.predicate = copy_expr(&s.predicate);
    n.aggs = copy_aggs(s.aggs, s.n_aggs);
    if (s.n_group_by > 0 && !s.group_by) fail(SQLRS_ERR_INVALID_ARG, "group_by is NULL");
    for (int32_t g = 0; g < s.n_group_by; g++) {
      n.group_by.push_back(copy_expr(&s.group_by[g]));
      n.group_names.push_back(s.group_names && s.group_names[g] ? s.group_names[g] : "");
    }
    if (s.kind == SQLRS_NODE_HASH_JOIN) {
      if (s.n_keys < 1) fail(SQLRS_ERR_INTERNAL, "HashJoin must has on condition");
      for (int32_t q = 0; q < s.n_keys; q++) {
        n.left_keys.push_back(copy_expr(&s.left_keys[q]));
        n.right_keys.push_back(copy_expr(&s.right_keys[q]));
      }
      n.join_fields = import_fields(s.join_output_schema);
    }
    auto check_child = [&](int c) {
      if (c < 0 || c >= n_nodes) fail(SQLRS_ERR_INVALID_ARG, "plan: child index out of range");
    };
    switch (s.kind) {
      case SQLRS_NODE_SCAN: break;
      case SQLRS_NODE_FILTER:
      case SQLRS_NODE_SIMPLE_AGG:
      case SQLRS_NODE_HASH_AGG: check_child(s.child0); break;
      case SQLRS_NODE_HASH_JOIN:
        check_child(s.child0);
        check_child(s.child1);
        break;
      default: fail(SQLRS_ERR_INVALID_ARG, "plan: unknown node kind");
    }
    nodes_.push_back(std::move(n));
  }
}

void Plan::push_table(int slot, DBatch batch) { tables_[slot].push_back(std::move(batch)); }

void Plan::reset() {
  for (Result& r : results_) {
    if (r.on_host) {
      if (r.arr.release) r.arr.release(&r.arr);
      if (r.sch.release) r.sch.release(&r.sch);
    }
  }
  results_.clear();
  tables_.clear();
  partial_active_ = false;
}

// aggregate at `idx` -> host Arrow.  A Filter directly below is fused into the aggregate's row
// program unless SQLRS_FLAG_NO_FUSION asks for operator-at-a-time execution.
void Plan::run_agg_to_host(int idx, Result* res) {
  Node& n = nodes_[idx];
  const bool simple = n.kind == SQLRS_NODE_SIMPLE_AGG;
  int child = n.child0;
  ExprCopy fused;
  if (!(opt_.flags & SQLRS_FLAG_NO_FUSION) && nodes_[child].kind == SQLRS_NODE_FILTER) {
    fused = nodes_[child].predicate;
    child = nodes_[child].child0;
    description_ += "[Filter+" + std::string(simple ? "SimpleAgg" : "HashAgg") + " fused] ";
  } else {
    description_ += std::string(simple ? "[SimpleAgg] " : "[HashAgg] ");
  }
  // the operator (compiled kernels, group table, scratch) is kept across execute() calls
  if (!agg_op_) agg_op_ = std::make_unique<AggOp>(n.aggs, n.group_by, n.group_names, simple, fused, opt_);
  else agg_op_->reset();
  AggOp& op = *agg_op_;
  for (const DBatch& b : run(child)) op.push(b);
  op.finish_host(&res->arr, &res->sch);
  res->on_host = true;
  description_ += op.describe() + "; ";
  scan_kernel_ms_ = op.scan_kernel_ms();
  scan_kernel_launches_ = op.scan_kernel_launches();
}

// ---- partial / final split for multi-GPU group-by: run everything below the root
// aggregate on this rank's shard and keep the group table un-finalised
void Plan::execute_partial(int64_t row_base) {
  description_.clear();
  ctx_.reap();
  Node& n = nodes_[root_];
  if (n.kind != SQLRS_NODE_SIMPLE_AGG && n.kind != SQLRS_NODE_HASH_AGG)
    fail(SQLRS_ERR_INVALID_ARG, "execute_partial: the plan root must be an aggregate");
  const bool simple = n.kind == SQLRS_NODE_SIMPLE_AGG;
  int child = n.child0;
  ExprCopy fused;
  if (!(opt_.flags & SQLRS_FLAG_NO_FUSION) && nodes_[child].kind == SQLRS_NODE_FILTER) {
    fused = nodes_[child].predicate;
    child = nodes_[child].child0;
    description_ += "[Filter+" + std::string(simple ? "SimpleAgg" : "HashAgg") + " fused, partial] ";
  }
  if (!partial_op_) partial_op_ = std::make_unique<AggOp>(n.aggs, n.group_by, n.group_names, simple, fused, opt_);
  else partial_op_->reset();
  partial_active_ = true;
  partial_op_->set_row_base(row_base);
  for (const DBatch& b : run(child)) partial_op_->push(b);
  description_ += partial_op_->describe() + "; ";
  scan_kernel_ms_ = partial_op_->scan_kernel_ms();
  scan_kernel_launches_ = partial_op_->scan_kernel_launches();
}
int Plan::partial_row_words() const {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "partial_row_words before execute_partial");
  return partial_op_->partial_row_words();
}
void Plan::export_partials_device(uint64_t* dst, int64_t cap_rows) {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "export_partials before execute_partial");
  partial_op_->export_partials_device(dst, cap_rows);
}
void Plan::merge_partials_device(const uint64_t* src, int n_bufs, int64_t cap_rows) {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "merge_partials before execute_partial");
  partial_op_->merge_partials_device(src, n_bufs, cap_rows);
}
void Plan::export_partials(ArrowArray* out, ArrowSchema* out_schema) {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "export_partials before execute_partial");
  partial_op_->export_partials(out, out_schema);
}
void Plan::clear_partials() {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "clear_partials before execute_partial");
  partial_op_->clear_partials();
}
void Plan::merge_partials(const DBatch& partials) {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "merge_partials before execute_partial");
  partial_op_->merge_partials(partials);
}
void Plan::finish_partial() {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "finish before execute_partial");
  results_.emplace_back();
  partial_op_->finish_host(&results_.back().arr, &results_.back().sch);
  results_.back().on_host = true;
  partial_active_ = false;
}

std::vector<DBatch> Plan::run(int idx) {
  Node& n = nodes_[idx];
  switch (n.kind) {
    case SQLRS_NODE_SCAN: return tables_[n.table_slot];
    case SQLRS_NODE_FILTER: {
      if (!n.filter_prog) {
        EvalRequest r;
        r.exprs.push_back(n.predicate);
        r.is_key.push_back(false);
        r.outs.push_back({OUT_KEEP, 0});
        n.filter_prog = std::make_unique<EvalProgram>(std::move(r));
      }
      description_ += "[Filter: sq_eval_kernel keep-mask + ballot compaction + gather] ";
      std::vector<DBatch> out;
      for (const DBatch& b : run(n.child0)) out.push_back(filter_batch(ctx_, *n.filter_prog, b));
      return out;
    }
    case SQLRS_NODE_SIMPLE_AGG:
    case SQLRS_NODE_HASH_AGG:
      fail(SQLRS_ERR_UNSUPPORTED, "an aggregate below another operator is not supported by the CUDA plan executor yet");
    case SQLRS_NODE_HASH_JOIN: {
      description_ += "[HashJoin: hash build (CSR) + probe count/scan/write + gathers] ";
      JoinOp j(n.join_type, n.left_keys, n.right_keys, n.predicate, n.join_fields, opt_);
      for (const DBatch& b : run(n.child0)) j.build_push(b);
      std::vector<DBatch> out;
      for (const DBatch& b : run(n.child1)) {
        DBatch r;
        if (j.probe(b, &r)) out.push_back(r);
      }
      DBatch tail;
      if (j.finish(&tail)) out.push_back(tail);
      return out;
    }
  }
  fail(SQLRS_ERR_INVALID_ARG, "plan: unknown node kind");
}

void Plan::execute() {
  for (Result& r : results_) {
    if (r.on_host) {
      if (r.arr.release) r.arr.release(&r.arr);
      if (r.sch.release) r.sch.release(&r.sch);
    }
  }
  results_.clear();
  description_.clear();
  ctx_.reap();
  Node& root = nodes_[root_];
  if (root.kind == SQLRS_NODE_SIMPLE_AGG || root.kind == SQLRS_NODE_HASH_AGG) {
    results_.emplace_back();
    run_agg_to_host(root_, &results_.back());
    return;
  }
  for (DBatch& b : run(root_)) {
    results_.emplace_back();
    results_.back().dev = std::move(b);
  }
}

bool Plan::next(ArrowArray* out, ArrowSchema* out_schema) {
  if (results_.empty()) return false;
  Result& r = results_.front();
  if (r.on_host) {
    *out = r.arr;
    if (out_schema) *out_schema = r.sch;
    else if (r.sch.release) r.sch.release(&r.sch);
  } else {
    export_batch_host(ctx_, r.dev, out, out_schema);
  }
  results_.pop_front();
  return true;
}

}  // namespace sq
