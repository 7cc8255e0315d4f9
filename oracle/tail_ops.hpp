// ORACLE — TEST INFRASTRUCTURE ONLY.
// CPU restatement of the operators that FOLLOW the hot path in a v1 plan:
//   ProjectExecutor  src/executor/project.rs:6-29
//   OrderExecutor    src/executor/order.rs:8-67   (arrow 28 lexsort_to_indices / sort_to_indices + take)
//   LimitExecutor    src/executor/limit.rs:6-80
//   CrossJoinExecutor src/executor/join/cross_join.rs:8-57 (pinned by tests/slt/join.slt:96-103)
// Plan position in the reference: Agg -> Order -> Project -> Limit (src/planner/select.rs:34-45).
//
// Pinned by the reference's goldens: limit.rs:96-101 (six offset/limit cases over 1 and 3 batches),
// executor/mod.rs:353-396, tests/slt/order.slt, tests/slt/limit.slt (tests/test_reference_golden.py).
// UNPINNED: the order of rows that compare EQUAL on every sort key — arrow sorts the index vector with
// `sort_unstable_by`, so the reference itself does not define it; this restatement (and the CUDA
// library) keep ties in input order.  Arrow semantics restated from arrow-rs 28.0.0 (not vendored under
// /root/reference): SortOptions::default() = ascending, nulls first; `descending` reverses the value
// comparison only (NULLs stay first); a single sort column goes through sort_to_indices, which in descending
// mode also reverses the run of NULL rows; floats compare by total order (f64::total_cmp).
#pragma once
#include <numeric>

#include "ops.hpp"

namespace oracle {

// ------------------------------------------------------------------ Project
// project.rs:14-28: one output batch per input batch; columns = eval_column of every expression; the field of
// an InputRef is the input field itself (evaluator.rs:31), every other expression yields a nullable field
// whose name the host computes (evaluator.rs:33-63).
struct Project {
  std::vector<Expr> exprs;
  std::vector<std::string> names;
  std::vector<uint8_t> keep_field;  // names[k] was NULL: a bare InputRef keeps the input field
  Batch execute(const Batch& in) const {
    Batch out;
    out.n = in.n;
    for (size_t k = 0; k < exprs.size(); k++) {
      ColPtr c = eval_expr(exprs[k], in);
      out.cols.push_back(c);
      const ExprNode& root = exprs[k].back();
      if (exprs[k].size() == 1 && root.op == SQLRS_OP_INPUT_REF && k < keep_field.size() && keep_field[k]) out.fields.push_back(in.fields.at(root.index));
      else out.fields.push_back(Field{k < names.size() ? names[k] : "", c->dtype, true});
    }
    return out;
  }
};

// ------------------------------------------------------------------ Order
inline int64_t f64_total_order_key(double d) {  // f64::total_cmp as a signed integer comparison
  int64_t b;
  std::memcpy(&b, &d, 8);
  return b ^ (int64_t)((uint64_t)(b >> 63) >> 1);
}
// -1 / 0 / +1 for two VALID cells of one column
inline int compare_cells(const Column& c, int64_t a, int64_t b) {
  switch (c.dtype) {
    case SQLRS_DT_FLOAT64: {
      const int64_t x = f64_total_order_key(c.f[a]), y = f64_total_order_key(c.f[b]);
      return x < y ? -1 : (x > y ? 1 : 0);
    }
    case SQLRS_DT_UTF8: {
      const int r = c.s[a].compare(c.s[b]);
      return r < 0 ? -1 : (r > 0 ? 1 : 0);
    }
    default: return c.i[a] < c.i[b] ? -1 : (c.i[a] > c.i[b] ? 1 : 0);
  }
}

struct Order {
  std::vector<Expr> exprs;
  std::vector<uint8_t> asc;
  std::vector<Batch> batches;

  void push(const Batch& b) { batches.push_back(b); }

  // order.rs:26-66
  Batch finish() {
    if (batches.empty()) fail(SQLRS_ERR_INTERNAL, "called `Option::unwrap()` on a `None` value (no input batch)");  // :27
    Batch all;
    all.fields = batches[0].fields;
    for (size_t c = 0; c < all.fields.size(); c++) {
      std::vector<ColPtr> parts;
      for (const Batch& b : batches) {
        if (b.cols.size() != all.fields.size()) fail(SQLRS_ERR_ARROW, "concat_batches: schema mismatch");
        parts.push_back(b.cols[c]);
      }
      all.cols.push_back(concat_columns(parts));
    }
    for (const Batch& b : batches) all.n += b.n;
    std::vector<ColPtr> keys;
    for (const Expr& e : exprs) {
      ColPtr k = eval_expr(e, all);
      keys.push_back(k);
    }
    std::vector<int64_t> idx((size_t)all.n);
    std::iota(idx.begin(), idx.end(), 0);
    if (keys.size() == 1) {
      // sort_to_indices: NULLs first (reversed when descending), then the valid rows by value
      const Column& k = *keys[0];
      std::vector<int64_t> nulls, valids;
      for (int64_t r = 0; r < all.n; r++) (k.is_valid(r) ? valids : nulls).push_back(r);
      const bool desc = !asc[0];
      std::stable_sort(valids.begin(), valids.end(), [&](int64_t a, int64_t b) {
        const int c = compare_cells(k, a, b);
        return desc ? c > 0 : c < 0;
      });
      if (desc) std::reverse(nulls.begin(), nulls.end());
      idx = nulls;
      idx.insert(idx.end(), valids.begin(), valids.end());
    } else if (keys.size() > 1) {
      // LexicographicalComparator: (NULL, NULL) equal -> next column; NULL before any value (nulls_first)
      std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) {
        for (size_t c = 0; c < keys.size(); c++) {
          const Column& k = *keys[c];
          const bool va = k.is_valid(a), vb = k.is_valid(b);
          if (!va && !vb) continue;
          if (!va) return true;
          if (!vb) return false;
          int r = compare_cells(k, a, b);
          if (!asc[c]) r = -r;
          if (r != 0) return r < 0;
        }
        return false;
      });
    }
    Batch out;
    out.fields = all.fields;
    out.n = all.n;
    for (const ColPtr& c : all.cols) out.cols.push_back(take(*c, idx));
    return out;
  }
};

// ------------------------------------------------------------------ CrossJoin
// cross_join.rs:26-56: the left side is drained and concatenated; for every right batch and every left row one output
// batch = that row's scalars materialised to the right batch's length (build_scalar_value_array) + the right columns.
struct CrossJoin {
  std::vector<Field> out_fields;
  std::vector<Batch> left_batches;
  bool sealed = false;
  Batch left_single;

  void build_push(const Batch& b) {
    if (sealed) fail(SQLRS_ERR_INVALID_ARG, "cross_join: build_push after probe");
    left_batches.push_back(b);
  }
  void seal() {
    if (sealed) return;
    sealed = true;
    if (left_batches.empty()) return;
    left_single.fields = left_batches[0].fields;
    for (size_t c = 0; c < left_single.fields.size(); c++) {
      std::vector<ColPtr> parts;
      for (const Batch& b : left_batches) {
        if (b.cols.size() != left_single.fields.size()) fail(SQLRS_ERR_ARROW, "concat_batches: schema mismatch");
        parts.push_back(b.cols[c]);
      }
      left_single.cols.push_back(concat_columns(parts));
    }
    for (const Batch& b : left_batches) left_single.n += b.n;
  }
  std::vector<Batch> probe(const Batch& right) {
    seal();
    std::vector<Batch> out;
    if (left_batches.empty()) return out;  // :33-35
    if (left_single.cols.size() + right.cols.size() != out_fields.size())
      fail(SQLRS_ERR_ARROW, "number of columns must match number of fields in schema");
    for (int64_t r = 0; r < left_single.n; r++) {
      Batch b;
      b.fields = out_fields;
      b.n = right.n;
      for (const ColPtr& c : left_single.cols) b.cols.push_back(scalar_to_column(Scalar::from_column(*c, r), right.n));
      for (const ColPtr& c : right.cols) b.cols.push_back(c);
      for (size_t k = 0; k < b.cols.size(); k++)
        if (b.cols[k]->dtype != out_fields[k].dtype)
          fail(SQLRS_ERR_ARROW, std::string("column types must match schema types, expected ") + dtype_name(out_fields[k].dtype) + " but found " +
                                    dtype_name(b.cols[k]->dtype));
      out.push_back(std::move(b));
    }
    return out;
  }
};

// ------------------------------------------------------------------ Limit
inline ColPtr slice_column(const Column& src, int64_t start, int64_t len) {
  std::vector<int64_t> idx((size_t)len);
  std::iota(idx.begin(), idx.end(), start);
  return take(src, idx);
}

// limit.rs:14-79, variable for variable.  NOTE (reference behaviour, unpinned): with `limit: None` the per-batch
// `limit_val` is the CURRENT batch's row count, so `OFFSET k` without LIMIT over several batches stops at
// offset + (rows of the batch being looked at).
struct Limit {
  int64_t limit = -1, offset = -1;  // -1 = None
  int64_t returned_count = 0;
  bool done = false;

  // returns true when a batch is yielded
  bool push(const Batch& batch, Batch* out) {
    if (limit == 0) {  // :31-33 `return Ok(())`
      done = true;
      return false;
    }
    if (done) return false;
    const int64_t offset_val = offset < 0 ? 0 : offset;
    const int64_t cardinality = batch.n;
    const int64_t limit_val = limit < 0 ? cardinality : limit;
    const int64_t start = std::max(returned_count, offset_val) - returned_count;
    const int64_t total_end = offset_val + limit_val;
    const int64_t current_batch_end = returned_count + cardinality;
    const int64_t real_end = std::min(total_end, current_batch_end);
    if (real_end < returned_count)  // usize underflow at :58 (only reachable with limit None over several batches)
      fail(SQLRS_ERR_INTERNAL, "attempt to subtract with overflow (limit.rs:58)");
    const int64_t end = real_end - returned_count;
    returned_count += cardinality;
    if (start >= end) return false;  // :63-65 `continue` (the break test below is skipped, as in the reference)
    if (start == 0 && end == cardinality) {
      *out = batch;
    } else {
      out->fields = batch.fields;
      out->n = end - start;
      out->cols.clear();
      for (const ColPtr& c : batch.cols) out->cols.push_back(slice_column(*c, start, end - start));
    }
    if (returned_count >= offset_val + limit_val) done = true;  // :76-78
    return true;
  }
};

}  // namespace oracle
