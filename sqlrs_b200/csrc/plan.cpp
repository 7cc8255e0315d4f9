// sqlrs_b200 — physical sub-plan executor (see plan.hpp).
#include "plan.hpp"

namespace sq {

// ---- which input columns does an expression read?
static void mark_refs(const ExprCopy& e, std::vector<bool>& needed, int offset = 0, int lo = 0, int hi = 1 << 30) {
  for (const ExprNodeCopy& n : e)
    if (n.op == SQLRS_OP_INPUT_REF && n.index >= lo && n.index < hi) {
      const int k = n.index - offset;
      if (k >= 0) {
        if (k >= (int)needed.size()) needed.resize(k + 1, false);
        needed[k] = true;
      }
    }
}


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
    if (s.kind == SQLRS_NODE_CROSS_JOIN) n.join_fields = import_fields(s.join_output_schema);
    if (s.kind == SQLRS_NODE_PROJECT || s.kind == SQLRS_NODE_ORDER) {
      if (s.n_exprs < 1 || !s.exprs) fail(SQLRS_ERR_INVALID_ARG, "plan: Project / Order need at least one expression");
      for (int32_t q = 0; q < s.n_exprs; q++) {
        n.exprs.push_back(copy_expr(&s.exprs[q]));
        n.expr_names.push_back(s.expr_names && s.expr_names[q] ? s.expr_names[q] : "");
        n.asc.push_back(s.order_asc ? s.order_asc[q] != 0 : true);
        n.keep_field.push_back(!(s.expr_names && s.expr_names[q]));
      }
    }
    if (s.kind == SQLRS_NODE_LIMIT) {
      n.limit = s.limit < 0 ? -1 : s.limit;
      n.offset = s.offset < 0 ? -1 : s.offset;
    }
    auto check_child = [&](int c) {
      if (c < 0 || c >= n_nodes) fail(SQLRS_ERR_INVALID_ARG, "plan: child index out of range");
    };
    switch (s.kind) {
      case SQLRS_NODE_SCAN: break;
      case SQLRS_NODE_FILTER:
      case SQLRS_NODE_SIMPLE_AGG:
      case SQLRS_NODE_HASH_AGG:
      case SQLRS_NODE_PROJECT:
      case SQLRS_NODE_ORDER:
      case SQLRS_NODE_LIMIT: check_child(s.child0); break;
      case SQLRS_NODE_HASH_JOIN:
      case SQLRS_NODE_CROSS_JOIN:
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
void Plan::clear_table(int slot) { tables_.erase(slot); }

// aggregate at `idx`, up to (excluding) finalisation.  A Filter directly below is fused into the aggregate's row
// program unless SQLRS_FLAG_NO_FUSION asks for operator-at-a-time execution.
AggOp& Plan::run_agg(int idx) {
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
  if (!n.agg_op) n.agg_op = std::make_unique<AggOp>(n.aggs, n.group_by, n.group_names, simple, fused, opt_);
  else n.agg_op->reset();
  AggOp& op = *n.agg_op;
  const Needed need = fusion() ? agg_child_needs(n, fused, width_of(child)) : Needed();
  if (!feed_fused_join(op, child, fused, need))
    for (const DBatch& b : run(child, need)) op.push(b);
  description_ += op.describe() + "; ";
  scan_kernel_ms_ = op.scan_kernel_ms();
  scan_kernel_launches_ = op.scan_kernel_launches();
  scan_from_partial_ = false;
  return op;
}

void Plan::run_agg_to_host(int idx, Result* res) {
  run_agg(idx).finish_host(&res->arr, &res->sch);
  res->on_host = true;
}

// An INNER HashJoin directly below the aggregate (no Filter in between, SQL key comparison): build the join's left
// side as usual, then probe and aggregate in ONE kernel per probe batch (csrc/jit/joinagg.cuh) — the joined rows
// are never materialised.  Returns false when the shape does not apply (the caller runs operator at a time).
bool Plan::feed_fused_join(AggOp& op, int child, const ExprCopy& agg_fused_pred, const Needed& need) {
  if (!fusion() || !agg_fused_pred.empty() || op.has_distinct()) return false;
  Node& jn = nodes_[child];
  if (jn.kind != SQLRS_NODE_HASH_JOIN || jn.join_type != SQLRS_JOIN_INNER || opt_.match_mode != SQLRS_MATCH_HASH_AND_KEY) return false;
  int left = jn.child0, right = jn.child1;
  ExprCopy build_pred, probe_pred;
  if (nodes_[left].kind == SQLRS_NODE_FILTER) {
    build_pred = nodes_[left].predicate;
    left = nodes_[left].child0;
  }
  if (nodes_[right].kind == SQLRS_NODE_FILTER) {
    probe_pred = nodes_[right].predicate;
    right = nodes_[right].child0;
  }
  const int nleft = width_of(left), total = (int)jn.join_fields.size();
  if (nleft <= 0 || nleft > total) return false;
  Needed left_need((size_t)nleft, false), right_need((size_t)(total - nleft), false);
  for (int k = 0; k < total; k++)
    if (need.empty() || ((size_t)k < need.size() && need[(size_t)k])) (k < nleft ? left_need[(size_t)k] : right_need[(size_t)(k - nleft)]) = true;
  mark_refs(jn.predicate, left_need, 0, 0, nleft);
  mark_refs(jn.predicate, right_need, nleft, nleft, 1 << 30);
  for (const ExprCopy& e : jn.left_keys) mark_refs(e, left_need);
  for (const ExprCopy& e : jn.right_keys) mark_refs(e, right_need);
  mark_refs(build_pred, left_need);
  mark_refs(probe_pred, right_need);
  JoinOp j(jn.join_type, jn.left_keys, jn.right_keys, jn.predicate, jn.join_fields, opt_);
  j.set_side_predicates(build_pred, ExprCopy());
  j.enable_fused_build();
  bool chained = false;
  if (build_pred.empty() && nodes_[left].kind == SQLRS_NODE_HASH_JOIN) chained = try_chain(child, left, left_need, j);
  if (chained) {
    description_ += "[HashJoin 1 probe builds HashJoin 2's table: sq_joinchain_kernel | probe 2 fused into the aggregate: sq_joinagg_kernel] ";
  } else {
    description_ += "[HashJoin build (CSR) | probe fused into the aggregate: sq_joinagg_kernel] ";
    for (const DBatch& b : run(left, left_need)) j.build_push(b);
  }
  const bool defer = chained && nodes_[child].chain_op->pending();
  for (const DBatch& b : run(right, right_need)) op.push_join(b, j, probe_pred, defer);
  return true;
}

// (A join B) join C with A join B inner, no non-equi filter, and C's join reading nothing of A: B's probe of A's table
// inserts (key 2 -> B row) straight into join 2's table; join 1's output is never materialised.
bool Plan::try_chain(int jidx, int left, const Needed& left_need, JoinOp& j) {
  Node& j2 = nodes_[jidx];
  Node& j1 = nodes_[left];
  if (j1.join_type != SQLRS_JOIN_INNER || !j1.predicate.empty() || j2.left_keys.size() != 1 || j1.left_keys.empty()) return false;
  int l1 = j1.child0, r1 = j1.child1;
  ExprCopy build_pred1, probe_pred1;
  if (nodes_[l1].kind == SQLRS_NODE_FILTER) {
    build_pred1 = nodes_[l1].predicate;
    l1 = nodes_[l1].child0;
  }
  if (nodes_[r1].kind == SQLRS_NODE_FILTER) {
    probe_pred1 = nodes_[r1].predicate;
    r1 = nodes_[r1].child0;
  }
  const int nleft1 = width_of(l1), total1 = (int)j1.join_fields.size();
  if (nleft1 <= 0 || nleft1 > total1) return false;
  for (int k = 0; k < nleft1; k++)
    if (left_need.empty() || ((size_t)k < left_need.size() && left_need[(size_t)k])) return false;  // a build-1 column is read above
  if (nodes_[r1].kind != SQLRS_NODE_SCAN) return false;  // the probe side must be a resident table (ONE batch, checked below)
  auto it = tables_.find(nodes_[r1].table_slot);
  if (it == tables_.end() || it->second.size() != 1 || it->second[0].n <= 0) return false;
  if (!j2.chain_op) j2.chain_op = std::make_unique<JoinChainOp>(opt_);
  if (j2.chain_op->disabled()) return false;
  Needed l1_need((size_t)nleft1, false);
  for (const ExprCopy& e : j1.left_keys) mark_refs(e, l1_need);
  mark_refs(build_pred1, l1_need);
  JoinOp jo1(j1.join_type, j1.left_keys, j1.right_keys, j1.predicate, j1.join_fields, opt_);
  jo1.set_side_predicates(build_pred1, ExprCopy());
  jo1.enable_fused_build();
  for (const DBatch& b : run(l1, l1_need)) jo1.build_push(b);
  if (!j2.chain_op->run(jo1, it->second[0], probe_pred1, j2.left_keys[0], j)) return false;
  if (j2.chain_op->pending()) pending_chains_.push_back(j2.chain_op.get());
  return true;
}

void Plan::run_validated(const std::function<void()>& body) {
  for (int attempt = 0;; attempt++) {
    pending_chains_.clear();
    bool ok = true;
    try {
      body();
    } catch (const RetrySizingError&) {
      ok = false;
    }
    if (ok && !pending_chains_.empty()) {
      SQ_CUDA(cudaStreamSynchronize(ctx_.stream));
      for (JoinChainOp* c : pending_chains_) ok = c->validate() && ok;
    } else {
      for (JoinChainOp* c : pending_chains_) c->validate();  // settles their state (the failed run's numbers are not used)
    }
    pending_chains_.clear();
    if (ok) return;
    if (std::getenv("SQLRS_B200_LOG_RETRY")) fprintf(stderr, "[sqlrs] plan run %d discarded: a hint-sized structure did not validate; re-running with exact sizing\n", attempt);
    if (attempt >= 3) fail(SQLRS_ERR_INTERNAL, "plan: sizing hints kept failing");
    description_.clear();
  }
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
  const std::string head = description_;
  run_validated([&]() {
    description_ = head;
    partial_op_->reset();
    partial_active_ = true;
    partial_sel_ = 0;
    partial_row_base_ = row_base;
    partial_op_->set_defer_tier_check(fusion() && partial_defer_ok_);
    partial_op_->set_row_base(row_base);
    const Needed need = fusion() ? agg_child_needs(n, fused, width_of(child)) : Needed();
    if (!feed_fused_join(*partial_op_, child, fused, need))
      for (const DBatch& b : run(child, need)) partial_op_->push(b);
    partial_op_->settle_hint_sized();  // a hint-sized group table is checked here (one counter read), before anything is exported
  });
  description_ += partial_op_->describe() + "; ";
  scan_from_partial_ = true;  // (read on demand: a deferred launch's timer is not waited for here)
}

// a partial run whose first sq_agg_small launch went unchecked (AggOp::set_defer_tier_check): read its status now; if it ran out
// of slots, run the partial aggregation again the checked way (the tier escalates inside that run)
void Plan::settle_partial() {
  if (!partial_op_ || !partial_op_->tier_check_pending()) return;
  try {
    partial_op_->settle();
  } catch (const RetrySizingError&) {
    const int sel = partial_sel_;
    partial_defer_ok_ = false;
    execute_partial(partial_row_base_);
    partial_defer_ok_ = true;  // (the operator itself remembers that deferring does not pay for this input)
    partial_sel_ = sel;
  }
}
double Plan::scan_kernel_ms() {
  if (scan_from_partial_ && partial_op_) scan_kernel_ms_ = partial_op_->scan_kernel_ms();
  return scan_kernel_ms_;
}
int64_t Plan::scan_kernel_launches() {
  if (scan_from_partial_ && partial_op_) scan_kernel_launches_ = partial_op_->scan_kernel_launches();
  return scan_kernel_launches_;
}
int Plan::partials_tables() const {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "partials_tables before execute_partial");
  return partial_op_->partial_tables();
}
void Plan::select_partials_table(int index) {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "select_partials_table before execute_partial");
  if (index < 0 || index >= partial_op_->partial_tables()) fail(SQLRS_ERR_INVALID_ARG, "partials table index out of range");
  partial_sel_ = index;
}
int Plan::partial_row_words() const {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "partial_row_words before execute_partial");
  return partial_op_->partial_table(partial_sel_).partial_row_words();
}
void Plan::export_partials_device(uint64_t* dst, int64_t cap_rows) {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "export_partials before execute_partial");
  partial_op_->partial_table(partial_sel_).export_partials_device(dst, cap_rows);
  // a plan that owns its stream (options.stream == NULL) has no stream the caller could order against: the buffer is
  // complete when the call returns.  With a caller-provided stream the pack is ordered on that stream.
  if (ctx_.own_stream) SQ_CUDA(cudaStreamSynchronize(ctx_.stream));
}
int64_t Plan::export_partials_partitioned(uint64_t* dst, int n_parts, int64_t cap_rows) {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "export_partials before execute_partial");
  settle_partial();
  const int64_t g = partial_op_->partial_table(partial_sel_).export_partials_partitioned(dst, n_parts, cap_rows);
  if (ctx_.own_stream) SQ_CUDA(cudaStreamSynchronize(ctx_.stream));
  return g;
}
bool Plan::result_shape(int64_t* n_rows, int32_t* n_cols) {
  if (results_.empty()) return false;
  Result& r = results_.front();
  if (r.on_host) fail(SQLRS_ERR_UNSUPPORTED, "plan_result_shape: this result was finalised on the host (take sqlrs_plan_next)");
  if (n_rows) *n_rows = r.dev.n;
  if (n_cols) *n_cols = (int32_t)r.dev.cols.size();
  return true;
}
void Plan::next_to_device(void* const* columns, int32_t n_columns) {
  if (results_.empty()) fail(SQLRS_ERR_INVALID_ARG, "plan_next_to_device: no pending result");
  Result& r = results_.front();
  if (r.on_host) fail(SQLRS_ERR_UNSUPPORTED, "plan_next_to_device: this result was finalised on the host (take sqlrs_plan_next)");
  if (n_columns != (int32_t)r.dev.cols.size()) fail(SQLRS_ERR_INVALID_ARG, "plan_next_to_device: column count mismatch");
  for (int32_t c = 0; c < n_columns; c++) {
    DCol& col = r.dev.cols[(size_t)c];
    if (col.dtype == SQLRS_DT_NULL || col.dtype == SQLRS_DT_BOOL || col.dtype == SQLRS_DT_UTF8)
      fail(SQLRS_ERR_UNSUPPORTED, "plan_next_to_device: only fixed-width value columns (Int32 / Int64 / Float64)");
    if (col.valid && null_count_of(ctx_, col) > 0) fail(SQLRS_ERR_UNSUPPORTED, "plan_next_to_device: column with NULLs");
    if (!columns[c] && col.n > 0) fail(SQLRS_ERR_INVALID_ARG, "plan_next_to_device: NULL destination");
    if (col.n > 0)
      SQ_CUDA(cudaMemcpyAsync(columns[c], col.data, (size_t)col.n * dtype_width(col.dtype), cudaMemcpyDeviceToDevice, ctx_.stream));
  }
  DBatch keep = std::move(r.dev);
  results_.pop_front();
  ctx_.defer([keep]() {});
  if (ctx_.own_stream) SQ_CUDA(cudaStreamSynchronize(ctx_.stream));
}
// ONE imported batch scanned as zero-copy slices of batch_rows rows (pointer arithmetic; validity / Boolean words need a
// 32-row granularity)
void Plan::push_table_batched(int slot, const DBatch& whole, int64_t batch_rows) {
  if (batch_rows <= 0 || (batch_rows & 31)) fail(SQLRS_ERR_INVALID_ARG, "push_table_batched: batch_rows must be a positive multiple of 32");
  if (whole.n == 0) {
    tables_[slot].push_back(whole);
    return;
  }
  for (int64_t off = 0; off < whole.n; off += batch_rows) {
    DBatch b;
    b.fields = whole.fields;
    b.n = std::min(batch_rows, whole.n - off);
    for (const DCol& c : whole.cols) {
      DCol s = c;
      s.n = b.n;
      if (c.dtype == SQLRS_DT_NULL) s.null_count = b.n;
      else if (c.dtype == SQLRS_DT_BOOL) s.data = (const uint32_t*)c.data + (off >> 5);
      else s.data = (const uint8_t*)c.data + (size_t)off * dtype_width(c.dtype);
      if (c.valid) {
        s.valid = c.valid + (off >> 5);
        s.null_count = -1;
      }
      b.cols.push_back(s);
    }
    tables_[slot].push_back(std::move(b));
  }
}
void Plan::merge_partials_device(const uint64_t* src, int n_bufs, int64_t cap_rows) {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "merge_partials before execute_partial");
  settle_partial();
  partial_op_->partial_table(partial_sel_).merge_partials_device(src, n_bufs, cap_rows);
}
void Plan::export_partials(ArrowArray* out, ArrowSchema* out_schema) {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "export_partials before execute_partial");
  settle_partial();
  partial_op_->partial_table(partial_sel_).export_partials(out, out_schema);
}
void Plan::clear_partials() {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "clear_partials before execute_partial");
  partial_op_->partial_table(partial_sel_).clear_partials();
}
void Plan::merge_partials(const DBatch& partials) {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "merge_partials before execute_partial");
  settle_partial();
  partial_op_->partial_table(partial_sel_).merge_partials(partials);
}
void Plan::finish_partial() {
  if (!partial_active_) fail(SQLRS_ERR_INVALID_ARG, "finish before execute_partial");
  settle_partial();
  results_.emplace_back();
  partial_op_->finish_host(&results_.back().arr, &results_.back().sch);
  results_.back().on_host = true;
  partial_active_ = false;
}

int Plan::width_of(int idx) {
  const Node& n = nodes_[idx];
  switch (n.kind) {
    case SQLRS_NODE_SCAN: {
      auto it = tables_.find(n.table_slot);
      return it == tables_.end() || it->second.empty() ? 0 : (int)it->second[0].cols.size();
    }
    case SQLRS_NODE_FILTER:
    case SQLRS_NODE_ORDER:
    case SQLRS_NODE_LIMIT: return width_of(n.child0);
    case SQLRS_NODE_PROJECT: return (int)n.exprs.size();
    case SQLRS_NODE_HASH_JOIN:
    case SQLRS_NODE_CROSS_JOIN: return (int)n.join_fields.size();
    default: return (int)(n.group_by.size() + n.aggs.size());
  }
}

Plan::Needed Plan::agg_child_needs(const Node& agg, const ExprCopy& fused_pred, int child_width) const {
  Needed need((size_t)std::max(child_width, 0), false);
  for (const AggSpec& a : agg.aggs) mark_refs(a.arg, need);
  for (const ExprCopy& g : agg.group_by) mark_refs(g, need);
  mark_refs(fused_pred, need);
  return need;
}

// Pull-based like the reference's stream tree, but a whole child result at a time and entirely in HBM.
// `needed` prunes the columns materialising operators (Filter, HashJoin) gather: positions are kept, a
// pruned column is a Null-typed placeholder that nothing above references.
std::vector<DBatch> Plan::run(int idx, const Needed& needed) {
  Node& n = nodes_[idx];
  auto is_needed = [&](size_t k) { return needed.empty() || (k < needed.size() && needed[k]); };
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
      Needed child_need = needed;
      if (!child_need.empty()) mark_refs(n.predicate, child_need);
      std::vector<DBatch> out;
      for (const DBatch& b : run(n.child0, child_need)) {
        if (needed.empty()) {
          out.push_back(filter_batch(ctx_, *n.filter_prog, b));
        } else {  // gather only what is read above
          DBatch pruned = b;
          DBatch res = filter_batch(ctx_, *n.filter_prog, b, &needed);
          out.push_back(res);
        }
      }
      return out;
    }
    case SQLRS_NODE_SIMPLE_AGG:
    case SQLRS_NODE_HASH_AGG: {
      // an aggregate below another operator: its result stays in HBM (finalised by one kernel)
      DBatch r = run_agg(idx).finish_device();
      description_ += "[aggregate finalised on the device] ";
      return {r};
    }
    case SQLRS_NODE_CROSS_JOIN: {
      CrossJoinOp j(n.join_fields, opt_);
      description_ += "[CrossJoin: one batch per left row (fill + gather), right columns shared] ";
      for (const DBatch& b : run(n.child0, Needed())) j.build_push(b);
      std::vector<DBatch> out;
      for (const DBatch& b : run(n.child1, Needed()))
        for (DBatch& r : j.probe(ctx_, b)) out.push_back(std::move(r));
      return out;
    }
    case SQLRS_NODE_PROJECT: {
      if (!n.project_op) n.project_op = std::make_unique<ProjectOp>(n.exprs, n.expr_names, n.keep_field, opt_);
      Needed child_need;
      if (fusion()) {
        const int w = width_of(n.child0);
        if (w > 0) {
          child_need.assign((size_t)w, false);
          // ProjectOp evaluates EVERY expression, as the reference does (project.rs:20-24: an error in a select item
          // nobody reads still ends the stream), so every expression's inputs must arrive — not only those of the
          // columns the parent reads
          for (size_t k = 0; k < n.exprs.size(); k++) mark_refs(n.exprs[k], child_need);
        }
      }
      description_ += "[Project: sq_eval_kernel] ";
      std::vector<DBatch> out;
      for (const DBatch& b : run(n.child0, child_need)) out.push_back(n.project_op->execute(ctx_, b));
      return out;
    }
    case SQLRS_NODE_ORDER: {
      if (!n.order_op) n.order_op = std::make_unique<OrderOp>(n.exprs, n.asc, opt_);
      else n.order_op->reset();
      Needed child_need = needed;
      if (!child_need.empty())
        for (const ExprCopy& e : n.exprs) mark_refs(e, child_need);
      n.order_op->set_row_limit(fusion() ? n.row_limit_hint : -1);
      const bool topk = fusion() && n.order_op->topk_applies(n.row_limit_hint);
      description_ += topk ? "[Order + Limit: top-" + std::to_string(n.row_limit_hint) + " selection (k_topk_pass: warp-level, keys in registers), rows gathered] "
                           : (n.row_limit_hint >= 0 && fusion() ? "[Order + Limit: stable LSD radix sort of row ids, top-" + std::to_string(n.row_limit_hint) + " rows gathered] "
                                                                 : "[Order: stable LSD radix sort of row ids + gather] ");
      const int ck = nodes_[n.child0].kind;
      if (topk && (ck == SQLRS_NODE_SIMPLE_AGG || ck == SQLRS_NODE_HASH_AGG)) {
        // the aggregate's groups are not sorted by first appearance first: the selection breaks ties by the ordinals
        AggOp& agg = run_agg(n.child0);
        if (!agg.has_distinct()) {
          DCol first_row;
          DBatch r = agg.finish_device(&first_row);
          description_ += "[aggregate finalised on the device, unordered] ";
          n.order_op->set_tiebreak(first_row);
          n.order_op->push(r);
          return {n.order_op->finish(ctx_)};
        }
        n.order_op->push(agg.finish_device());
        return {n.order_op->finish(ctx_)};
      }
      for (const DBatch& b : run(n.child0, child_need)) n.order_op->push(b);
      return {n.order_op->finish(ctx_)};
    }
    case SQLRS_NODE_LIMIT: {
      std::vector<DBatch> out;
      LimitOp l(n.limit, n.offset, opt_);
      if (l.done()) return out;  // limit 0: the child is never polled (limit.rs:31-33)
      // top-k: an Order below (possibly under Projects, which are row-wise) yields ONE batch, of which only the
      // first offset + limit rows can reach the output
      int below = n.child0;
      while (nodes_[below].kind == SQLRS_NODE_PROJECT) below = nodes_[below].child0;
      if (nodes_[below].kind == SQLRS_NODE_ORDER) nodes_[below].row_limit_hint = l.rows_needed();
      std::vector<DBatch> in = run(n.child0, needed);
      if (nodes_[below].kind == SQLRS_NODE_ORDER) nodes_[below].row_limit_hint = -1;
      description_ += "[Limit] ";
      for (const DBatch& b : in) {
        DBatch r;
        if (l.push(ctx_, b, &r)) out.push_back(r);
        if (l.done()) break;
      }
      return out;
    }
    case SQLRS_NODE_HASH_JOIN: {
      JoinOp j(n.join_type, n.left_keys, n.right_keys, n.predicate, n.join_fields, opt_);
      int left = n.child0, right = n.child1;
      ExprCopy build_pred, probe_pred;
      if (fusion()) {  // Filters directly below the join run inside the join's key kernels
        if (nodes_[left].kind == SQLRS_NODE_FILTER) {
          build_pred = nodes_[left].predicate;
          left = nodes_[left].child0;
        }
        if (nodes_[right].kind == SQLRS_NODE_FILTER) {
          probe_pred = nodes_[right].predicate;
          right = nodes_[right].child0;
        }
      }
      const int nleft = width_of(left);
      Needed left_need, right_need;
      if (fusion() && nleft > 0) {
        const int total = (int)n.join_fields.size();
        left_need.assign((size_t)nleft, false);
        right_need.assign((size_t)std::max(total - nleft, 0), false);
        for (int k = 0; k < total; k++)
          if (is_needed((size_t)k)) (k < nleft ? left_need[(size_t)k] : right_need[(size_t)(k - nleft)]) = true;
        mark_refs(n.predicate, left_need, 0, 0, nleft);            // non-equi filter over the joined row
        mark_refs(n.predicate, right_need, nleft, nleft, 1 << 30);
        Needed out_need((size_t)total, false);                        // what the join itself must gather
        for (int k = 0; k < total; k++) out_need[(size_t)k] = k < nleft ? left_need[(size_t)k] : right_need[(size_t)(k - nleft)];
        j.set_needed_columns(out_need);
        for (const ExprCopy& e : n.left_keys) mark_refs(e, left_need);   // the children must still deliver key / predicate inputs
        for (const ExprCopy& e : n.right_keys) mark_refs(e, right_need);
        mark_refs(build_pred, left_need);
        mark_refs(probe_pred, right_need);
      }
      j.set_side_predicates(build_pred, probe_pred);
      if (fusion()) j.enable_fused_build();
      description_ += std::string("[HashJoin") + (build_pred.empty() && probe_pred.empty() ? "" : " + fused side Filter") +
                      ": CSR hash build, probe count/scan/write, pruned gathers] ";
      for (const DBatch& b : run(left, left_need)) j.build_push(b);
      std::vector<DBatch> out;
      for (const DBatch& b : run(right, right_need)) {
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
  ctx_.reap();
  run_validated([&]() {
    for (Result& r : results_) {  // a discarded attempt
      if (r.on_host) {
        if (r.arr.release) r.arr.release(&r.arr);
        if (r.sch.release) r.sch.release(&r.sch);
      }
    }
    results_.clear();
    description_.clear();
    Node& root = nodes_[root_];
    if (root.kind == SQLRS_NODE_SIMPLE_AGG || root.kind == SQLRS_NODE_HASH_AGG) {
      results_.emplace_back();
      run_agg_to_host(root_, &results_.back());
      return;
    }
    for (DBatch& b : run(root_, Needed())) {
      results_.emplace_back();
      results_.back().dev = std::move(b);
    }
  });
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
