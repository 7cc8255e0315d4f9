// sqlrs_b200 — Project / Order / Limit on device batches (see tail.hpp; kernels: csrc/jit/eval.cuh, kernels_sort.cu).
#include "tail.hpp"

#include <cstdlib>

#include "kernels_aot.hpp"

namespace sq {

// expressions -> (one EvalProgram over the non-trivial ones, slot per expression); a bare InputRef is passed through
// like the reference's Arc clone (evaluator.rs:15)
static std::unique_ptr<EvalProgram> value_program(const std::vector<ExprCopy>& exprs, std::vector<int>* slot) {
  EvalRequest r;
  slot->assign(exprs.size(), -1);
  for (size_t k = 0; k < exprs.size(); k++) {
    if (exprs[k].empty()) fail(SQLRS_ERR_INVALID_ARG, "empty expression");
    if (exprs[k].size() == 1 && exprs[k][0].op == SQLRS_OP_INPUT_REF) continue;
    (*slot)[k] = (int)r.exprs.size();
    r.outs.push_back({OUT_VALUE, (int)r.exprs.size()});
    r.exprs.push_back(exprs[k]);
    r.is_key.push_back(false);
  }
  if (r.exprs.empty()) return nullptr;
  return std::make_unique<EvalProgram>(std::move(r));
}

static std::vector<DCol> eval_values(Ctx& ctx, EvalProgram* prog, const std::vector<ExprCopy>& exprs, const std::vector<int>& slot,
                                     const DBatch& in, const char* what) {
  EvalResult res;
  if (prog) res = prog->run(ctx, in, what);
  std::vector<DCol> out;
  for (size_t k = 0; k < exprs.size(); k++) {
    if (slot[k] >= 0) {
      out.push_back(res.cols[(size_t)slot[k]]);
    } else {
      const int idx = exprs[k][0].index;
      if (idx < 0 || idx >= (int)in.cols.size()) fail(SQLRS_ERR_INTERNAL, "InputRef index out of bounds");
      out.push_back(in.cols[(size_t)idx]);
    }
  }
  return out;
}

// ------------------------------------------------------------------ Project, project.rs:14-28
ProjectOp::ProjectOp(std::vector<ExprCopy> exprs, std::vector<std::string> names, std::vector<bool> keep_field, const Options& opt)
    : ctx_(opt), exprs_(std::move(exprs)), names_(std::move(names)), keep_field_(std::move(keep_field)) {
  prog_ = value_program(exprs_, &slot_);
}
DBatch ProjectOp::execute(const DBatch& in) { return execute(ctx_, in); }
DBatch ProjectOp::execute(Ctx& ctx, const DBatch& in) {
  Trace tr("project.execute", ctx.stream);
  DBatch out;
  out.n = in.n;
  out.cols = eval_values(ctx, prog_.get(), exprs_, slot_, in, "projection");
  for (size_t k = 0; k < exprs_.size(); k++) {
    if (slot_[k] < 0 && k < keep_field_.size() && keep_field_[k]) {  // eval_field of an InputRef = the input field itself (evaluator.rs:31)
      const size_t idx = (size_t)exprs_[k][0].index;
      out.fields.push_back(idx < in.fields.size() ? in.fields[idx] : Field{"", out.cols[k].dtype, true});
    } else {
      out.fields.push_back(Field{k < names_.size() ? names_[k] : "", out.cols[k].dtype, true});
    }
  }
  return out;
}

// ------------------------------------------------------------------ helpers
DBatch concat_batches(Ctx& ctx, const std::vector<DBatch>& batches) {
  if (batches.size() == 1) return batches[0];
  DBatch all;
  all.fields = batches[0].fields;
  for (const DBatch& b : batches) all.n += b.n;
  for (size_t c = 0; c < batches[0].cols.size(); c++) {
    std::vector<DCol> parts;
    for (const DBatch& b : batches) {
      if (b.cols.size() != batches[0].cols.size()) fail(SQLRS_ERR_ARROW, "concat_batches: schema mismatch");
      parts.push_back(b.cols[c]);
    }
    all.cols.push_back(concat_cols(ctx, parts, batches[0].cols[c].dtype));
  }
  return all;
}

DBatch slice_batch(Ctx& ctx, const DBatch& in, int64_t start, int64_t len) {
  DBatch out;
  out.fields = in.fields;
  out.n = len;
  if (start + len >= (1LL << 32)) fail(SQLRS_ERR_UNSUPPORTED, "slice beyond 2^32 rows");
  BufPtr idx = dev_alloc(ctx, (size_t)std::max<int64_t>(len, 1) * 4);
  launch_iota_u32((uint32_t*)idx->p, len, (uint32_t)start, ctx.stream);
  for (const DCol& c : in.cols) out.cols.push_back(gather_col_u32(ctx, c, (const uint32_t*)idx->p, len));
  ctx.defer([idx]() {});  // the index list must outlive the enqueued gathers
  return out;
}

// ------------------------------------------------------------------ CrossJoin, cross_join.rs:26-56
CrossJoinOp::CrossJoinOp(std::vector<Field> out_fields, const Options& opt) : ctx_(opt), out_fields_(std::move(out_fields)) {}
void CrossJoinOp::build_push(const DBatch& b) {
  if (sealed_) fail(SQLRS_ERR_INVALID_ARG, "cross_join: build_push after probe");
  left_batches_.push_back(b);
}
std::vector<DBatch> CrossJoinOp::probe(const DBatch& right) { return probe(ctx_, right); }
std::vector<DBatch> CrossJoinOp::probe(Ctx& ctx, const DBatch& right) {
  Trace tr("cross_join.probe", ctx.stream);
  std::vector<DBatch> out;
  if (!sealed_) {
    sealed_ = true;
    has_left_ = !left_batches_.empty();
    if (has_left_) left_single_ = concat_batches(ctx, left_batches_);  // :37
    left_batches_.clear();
  }
  if (!has_left_) return out;  // :33-35
  if (left_single_.cols.size() + right.cols.size() != out_fields_.size())
    fail(SQLRS_ERR_ARROW, "number of columns must match number of fields in schema");
  const int64_t m = right.n;
  for (int64_t r = 0; r < left_single_.n; r++) {  // :44-55: one batch per left row
    BufPtr idx = dev_alloc(ctx, (size_t)std::max<int64_t>(m, 1) * 8);
    launch_fill_u64((uint64_t*)idx->p, m, (uint64_t)r, ctx.stream);
    DBatch b;
    b.fields = out_fields_;
    b.n = m;
    for (const DCol& c : left_single_.cols) b.cols.push_back(gather_col_i64(ctx, c, (const int64_t*)idx->p, m, false));  // build_scalar_value_array
    for (const DCol& c : right.cols) b.cols.push_back(c);
    for (size_t k = 0; k < b.cols.size(); k++)
      if (b.cols[k].dtype != out_fields_[k].dtype)
        fail(SQLRS_ERR_ARROW, std::string("column types must match schema types, expected ") + dtype_name(out_fields_[k].dtype) + " but found " +
                                  dtype_name(b.cols[k].dtype));
    ctx.defer([idx]() {});
    out.push_back(std::move(b));
  }
  return out;
}

// ------------------------------------------------------------------ Order, order.rs:14-66
OrderOp::OrderOp(std::vector<ExprCopy> order_by, std::vector<bool> asc, const Options& opt)
    : ctx_(opt), order_by_(std::move(order_by)), asc_(std::move(asc)), flags_(opt.flags) {
  if (asc_.size() != order_by_.size()) fail(SQLRS_ERR_INVALID_ARG, "order: one direction per sort expression");
  prog_ = value_program(order_by_, &slot_);
}
bool OrderOp::topk_applies(int64_t row_limit) const {
  return row_limit >= 1 && row_limit <= kTopKMaxRows && order_by_.size() <= (size_t)kTopKMaxKeys && !std::getenv("SQLRS_B200_NO_TOPK");
}
DBatch OrderOp::finish() { return finish(ctx_); }
DBatch OrderOp::finish(Ctx& ctx) {
  Trace tr("order.finish", ctx.stream);
  KernelEvent ev(flags_, ctx.stream, "order (sort keys + top-k / radix passes + gather)");
  if (batches_.empty()) fail(SQLRS_ERR_INTERNAL, "called `Option::unwrap()` on a `None` value (no input batch)");  // order.rs:27
  DBatch all = concat_batches(ctx, batches_);  // :28
  batches_.clear();
  const int64_t n = all.n;
  if (n >= (1LL << 31)) fail(SQLRS_ERR_UNSUPPORTED, "Order: more than 2^31 rows in one sort");
  std::vector<DCol> keys = eval_values(ctx, prog_.get(), order_by_, slot_, all, "order by");  // :30-43
  // Utf8 sort keys: the column holds string pool ids; sort their byte-wise ranks (arrow sorts Utf8 by the strings' bytes)
  BufPtr rank_table;
  for (DCol& k : keys) {
    if (k.dtype != SQLRS_DT_UTF8) continue;
    if (!rank_table) {
      const std::vector<int32_t> ranks = StringPool::instance().ranks();
      rank_table = dev_alloc(ctx, std::max<size_t>(ranks.size(), 1) * 4);
      if (!ranks.empty()) SQ_CUDA(cudaMemcpyAsync(rank_table->p, ranks.data(), ranks.size() * 4, cudaMemcpyHostToDevice, ctx.stream));
      SQ_CUDA(cudaStreamSynchronize(ctx.stream));  // `ranks` is a local vector
    }
    DCol r = make_col(ctx, SQLRS_DT_INT64, n, false);
    launch_str_rank((const int64_t*)k.data, n, (const int32_t*)rank_table->p, (int64_t*)col_data(r), ctx.stream);
    r.valid = k.valid;
    r.keep_valid = k.keep_valid;
    r.null_count = k.null_count;
    k = r;
  }
  // ORDER BY ... LIMIT k (the plan passed the Limit down): select the k first rows instead of sorting all n.  Not for arrow's
  // single-column descending sort over a column with NULLs, which also reverses the run of NULL rows (sort_pass).
  bool topk = topk_applies(row_limit_);
  for (size_t c = 0; c < keys.size(); c++)
    if (keys[c].dtype == SQLRS_DT_NULL || (keys.size() == 1 && !asc_[c] && keys[c].valid)) topk = false;
  if (topk) {
    TopKKeys tk{};
    tk.m = (int)keys.size();
    for (size_t c = 0; c < keys.size(); c++) {
      tk.dtype[c] = keys[c].dtype;
      tk.descending[c] = asc_[c] ? 0 : 1;
      tk.data[c] = keys[c].data;
      tk.valid[c] = keys[c].valid;
    }
    tk.tiebreak = has_tiebreak_ ? (const uint64_t*)tiebreak_.data : nullptr;
    const int64_t m = std::min(row_limit_, n);
    BufPtr perm = dev_alloc(ctx, (size_t)m * 4);
    launch_topk(tk, n, (int)m, (uint32_t*)perm->p, ctx.stream);
    DBatch out;
    out.fields = all.fields;
    out.n = m;
    for (const DCol& c : all.cols) out.cols.push_back(gather_col_u32(ctx, c, (const uint32_t*)perm->p, m));
    DCol tb = tiebreak_;
    ctx.defer([perm, keys, tb]() {});
    return out;
  }
  BufPtr perm = dev_alloc(ctx, (size_t)std::max<int64_t>(n, 1) * 4);
  launch_iota_u32((uint32_t*)perm->p, n, 0u, ctx.stream);
  // unordered input with first-appearance ordinals: they are the least significant sort key
  if (has_tiebreak_) sort_pass(SQLRS_DT_INT64, tiebreak_.data, nullptr, n, false, false, (uint32_t*)perm->p, ctx.stream);
  // lexsort_to_indices (:45) as an LSD sequence of stable passes, last sort expression first.  A single descending
  // column goes through arrow's sort_to_indices, which also reverses the run of NULL rows.
  const bool single = keys.size() == 1;
  for (size_t c = keys.size(); c-- > 0;)
    sort_pass(keys[c].dtype, keys[c].data, keys[c].valid, n, !asc_[c], single && !asc_[c], (uint32_t*)perm->p, ctx.stream);
  const int64_t m = row_limit_ >= 0 ? std::min(row_limit_, n) : n;
  DBatch out;
  out.fields = all.fields;
  out.n = m;
  for (const DCol& c : all.cols) out.cols.push_back(gather_col_u32(ctx, c, (const uint32_t*)perm->p, m));  // take, :47-63
  ctx.defer([perm, keys]() {});
  return out;
}

// ------------------------------------------------------------------ Limit, limit.rs:14-79 (variable for variable)
LimitOp::LimitOp(int64_t limit, int64_t offset, const Options& opt)
    : ctx_(opt), limit_(limit < 0 ? -1 : limit), offset_(offset < 0 ? -1 : offset) {
  done_ = limit_ == 0;  // :31-33
}
bool LimitOp::push(const DBatch& in, DBatch* out) { return push(ctx_, in, out); }
bool LimitOp::push(Ctx& ctx, const DBatch& batch, DBatch* out) {
  if (done_) return false;
  const int64_t offset_val = offset_ < 0 ? 0 : offset_;
  const int64_t cardinality = batch.n;
  const int64_t limit_val = limit_ < 0 ? cardinality : limit_;  // :40 — None means "this batch's row count"
  const int64_t start = std::max(returned_count_, offset_val) - returned_count_;
  const int64_t total_end = offset_val + limit_val;
  const int64_t current_batch_end = returned_count_ + cardinality;
  const int64_t real_end = std::min(total_end, current_batch_end);
  if (real_end < returned_count_) fail(SQLRS_ERR_INTERNAL, "attempt to subtract with overflow (limit.rs:58)");
  const int64_t end = real_end - returned_count_;
  returned_count_ += cardinality;
  if (start >= end) return false;  // :63-65
  if (start == 0 && end == cardinality) *out = batch;
  else *out = slice_batch(ctx, batch, start, end - start);
  if (returned_count_ >= offset_val + limit_val) done_ = true;  // :76-78
  return true;
}

}  // namespace sq
