// ORACLE — TEST INFRASTRUCTURE ONLY.
// Expression evaluation: BoundExpr::eval_column (src/executor/evaluator.rs:13-28) and
// binary_op (src/executor/array_compute.rs:70-90), one full-length temporary per node.
//
// arrow 28.0.0 kernel semantics relied on (crate not under /root/reference, Cargo.lock:43-44;
// restated from its documentation — SURVEY.md §8c):
//   add/subtract/multiply: no overflow detection, integers wrap; NULL in either side -> NULL
//   divide: integer division truncates; any valid zero divisor -> Err(DivideByZero)
//   *_dyn comparisons: NULL in either side -> NULL; operand types must be equal
//   and_kleene / or_kleene: three-valued logic
//   cast: numeric casts are "safe" (out of range -> NULL); Null -> T gives all-NULL
// Float comparison of NaN / -0.0 is UNPINNED (no reference test); IEEE semantics used.
#pragma once
#include <cmath>
#include <limits>

#include "columns.hpp"

namespace oracle {

struct ExprNode {
  int op = 0, dtype = 0, index = 0, is_null = 0;
  int64_t imm_bits = 0;
  std::string str;
};
using Expr = std::vector<ExprNode>;

inline Expr copy_expr(const sqlrs_expr* e) {
  Expr out;
  if (!e) return out;
  if (e->n_nodes > 0 && !e->nodes) fail(SQLRS_ERR_INVALID_ARG, "sqlrs_expr.nodes is NULL");
  for (int32_t k = 0; k < e->n_nodes; k++) {
    ExprNode n;
    n.op = e->nodes[k].op;
    n.dtype = e->nodes[k].dtype;
    n.index = e->nodes[k].index;
    n.is_null = e->nodes[k].is_null;
    n.imm_bits = e->nodes[k].imm_bits;
    if (e->nodes[k].str) n.str = e->nodes[k].str;
    out.push_back(n);
  }
  return out;
}

inline Scalar constant_scalar(const ExprNode& n) {
  Scalar v = Scalar::null_of(n.dtype);
  if (n.dtype == SQLRS_DT_NULL || n.is_null) return v;
  v.is_null = false;
  switch (n.dtype) {
    case SQLRS_DT_BOOL: v.i = n.imm_bits != 0; break;
    case SQLRS_DT_INT32: v.i = (int32_t)n.imm_bits; break;
    case SQLRS_DT_INT64: v.i = n.imm_bits; break;
    case SQLRS_DT_FLOAT64: std::memcpy(&v.f, &n.imm_bits, 8); break;
    case SQLRS_DT_UTF8: v.s = n.str; break;
  }
  return v;
}

inline bool is_numeric(int dt) { return dt == SQLRS_DT_INT32 || dt == SQLRS_DT_INT64 || dt == SQLRS_DT_FLOAT64; }

inline void merge_validity(Column& out, const Column& l, const Column& r) {
  if (l.valid.empty() && r.valid.empty()) return;
  out.valid.assign(out.n, 1);
  for (int64_t k = 0; k < out.n; k++) out.valid[k] = (uint8_t)(l.is_valid(k) && r.is_valid(k));
}

inline int64_t wrap_int(int dtype, int64_t v) { return dtype == SQLRS_DT_INT32 ? (int64_t)(int32_t)v : v; }

// arithmetic_op!, array_compute.rs:37-46
inline ColPtr arithmetic(const Column& l, const Column& r, int op_in) {
  // the v2 engine's *_checked kernels (src/function/scalar/arithmetic_function.rs:66-71,147-152,219-224,244-249): an integer
  // result that does not fit its type is ArrowError::ComputeError("Overflow happened on: ..."), Float64 is unchanged
  const bool checked = op_in >= SQLRS_OP_ADD_CHECKED && op_in <= SQLRS_OP_DIV_CHECKED;
  const int op = checked ? op_in - (SQLRS_OP_ADD_CHECKED - SQLRS_OP_ADD) : op_in;
  if (!is_numeric(l.dtype)) fail(SQLRS_ERR_UNSUPPORTED, "todo!: unsupported data type");  // :43
  if (r.dtype != l.dtype)
    fail(SQLRS_ERR_INTERNAL, "compute_op failed to downcast array");  // expect() at :23
  if (l.n != r.n) fail(SQLRS_ERR_ARROW, "Cannot perform math operation on arrays of different length");
  auto out = std::make_shared<Column>();
  out->alloc(l.dtype, l.n);
  merge_validity(*out, l, r);
  if (l.dtype == SQLRS_DT_FLOAT64) {
    for (int64_t k = 0; k < l.n; k++) {
      double a = l.f[k], b = r.f[k];
      switch (op) {
        case SQLRS_OP_ADD: out->f[k] = a + b; break;
        case SQLRS_OP_SUB: out->f[k] = a - b; break;
        case SQLRS_OP_MUL: out->f[k] = a * b; break;
        case SQLRS_OP_DIV:
          if (out->is_valid(k) && b == 0.0) fail(SQLRS_ERR_ARROW, "Divide by zero error");
          out->f[k] = a / b;
          break;
      }
    }
  } else {
    for (int64_t k = 0; k < l.n; k++) {
      uint64_t a = (uint64_t)l.i[k], b = (uint64_t)r.i[k];
      int64_t v = 0;
      switch (op) {
        case SQLRS_OP_ADD: v = (int64_t)(a + b); break;
        case SQLRS_OP_SUB: v = (int64_t)(a - b); break;
        case SQLRS_OP_MUL: v = (int64_t)(a * b); break;
        case SQLRS_OP_DIV:
          if (!out->is_valid(k)) break;
          if (r.i[k] == 0) fail(SQLRS_ERR_ARROW, "Divide by zero error");
          if (r.i[k] == -1) v = (int64_t)(0 - a);  // div_wrapping: MIN / -1 wraps
          else v = l.i[k] / r.i[k];
          break;
      }
      if (checked && out->is_valid(k)) {
        bool ovf = false;
        const int64_t x = l.i[k], y = r.i[k];
        if (l.dtype == SQLRS_DT_INT32) {
          const int64_t w = op == SQLRS_OP_ADD ? x + y : op == SQLRS_OP_SUB ? x - y : op == SQLRS_OP_MUL ? x * y : (y == -1 ? -x : 0);
          ovf = w != (int64_t)(int32_t)w;
        } else {
          int64_t w;
          switch (op) {
            case SQLRS_OP_ADD: ovf = __builtin_add_overflow(x, y, &w); break;
            case SQLRS_OP_SUB: ovf = __builtin_sub_overflow(x, y, &w); break;
            case SQLRS_OP_MUL: ovf = __builtin_mul_overflow(x, y, &w); break;
            default: ovf = x == INT64_MIN && y == -1; break;
          }
        }
        if (ovf) fail(SQLRS_ERR_ARROW, "Compute error: Overflow happened on: " + std::to_string(x) + " and " + std::to_string(y));
      }
      out->i[k] = wrap_int(l.dtype, v);
    }
  }
  return out;
}

template <typename T>
inline bool cmp_apply(int op, const T& a, const T& b) {
  switch (op) {
    case SQLRS_OP_GT: return a > b;
    case SQLRS_OP_LT: return a < b;
    case SQLRS_OP_GE: return a >= b;
    case SQLRS_OP_LE: return a <= b;
    case SQLRS_OP_EQ: return a == b;
    case SQLRS_OP_NE: return a != b;
  }
  return false;
}

// gt_dyn / lt_dyn / gt_eq_dyn / lt_eq_dyn / eq_dyn / neq_dyn, array_compute.rs:80-85
inline ColPtr comparison(const Column& l, const Column& r, int op) {
  if (l.dtype != r.dtype)
    fail(SQLRS_ERR_ARROW, std::string("Invalid argument error: comparing ") + dtype_name(l.dtype) + " with " +
                              dtype_name(r.dtype));
  if (l.dtype == SQLRS_DT_NULL) fail(SQLRS_ERR_ARROW, "comparison of Null arrays is not supported");
  if (l.n != r.n) fail(SQLRS_ERR_ARROW, "Cannot compare arrays of different lengths");
  auto out = std::make_shared<Column>();
  out->alloc(SQLRS_DT_BOOL, l.n);
  merge_validity(*out, l, r);
  for (int64_t k = 0; k < l.n; k++) {
    bool v;
    if (l.dtype == SQLRS_DT_FLOAT64) v = cmp_apply(op, l.f[k], r.f[k]);
    else if (l.dtype == SQLRS_DT_UTF8) v = cmp_apply(op, l.s[k], r.s[k]);
    else v = cmp_apply(op, l.i[k], r.i[k]);
    out->i[k] = v;
  }
  return out;
}

// boolean_op! + and_kleene / or_kleene, array_compute.rs:48-68,86-87
inline ColPtr kleene(const Column& l, const Column& r, int op) {
  if (l.dtype != SQLRS_DT_BOOL || r.dtype != SQLRS_DT_BOOL)
    fail(SQLRS_ERR_INTERNAL, std::string("Cannot evaluate binary expression with types ") + dtype_name(l.dtype) +
                                 " and " + dtype_name(r.dtype) + ", only Boolean supported");
  if (l.n != r.n) fail(SQLRS_ERR_ARROW, "Cannot perform bitwise operation on arrays of different length");
  auto out = std::make_shared<Column>();
  out->alloc(SQLRS_DT_BOOL, l.n);
  bool any_null = !l.valid.empty() || !r.valid.empty();
  if (any_null) out->valid.assign(l.n, 1);
  for (int64_t k = 0; k < l.n; k++) {
    bool lv = l.is_valid(k), rv = r.is_valid(k);
    bool a = l.i[k] != 0, b = r.i[k] != 0;
    bool res = false, valid = true;
    if (op == SQLRS_OP_AND) {
      if (lv && rv) res = a && b;
      else if ((lv && !a) || (rv && !b)) res = false;  // false AND NULL = false
      else valid = false;
    } else {
      if (lv && rv) res = a || b;
      else if ((lv && a) || (rv && b)) res = true;  // true OR NULL = true
      else valid = false;
    }
    out->i[k] = res;
    if (any_null) out->valid[k] = valid;
  }
  out->normalize();
  return out;
}

// arrow compute::cast (evaluator.rs:23, sum.rs:54), the subset reachable from the v1 types
inline ColPtr cast_column(const ColPtr& src, int to) {
  const Column& c = *src;
  if (c.dtype == to) return src;
  if (c.dtype == SQLRS_DT_NULL) return null_column(to, c.n);
  auto out = std::make_shared<Column>();
  out->alloc(to, c.n);
  out->valid = c.valid;
  auto from_int = [&](int64_t lo, int64_t hi) {
    for (int64_t k = 0; k < c.n; k++) {
      int64_t v = c.i[k];
      if (v < lo || v > hi) {
        out->set_null(k);
        v = 0;
      }
      out->i[k] = v;
    }
  };
  if ((c.dtype == SQLRS_DT_INT32 || c.dtype == SQLRS_DT_INT64 || c.dtype == SQLRS_DT_BOOL) && to == SQLRS_DT_INT64) {
    out->i = c.i;
  } else if ((c.dtype == SQLRS_DT_INT64 || c.dtype == SQLRS_DT_BOOL) && to == SQLRS_DT_INT32) {
    from_int(std::numeric_limits<int32_t>::min(), std::numeric_limits<int32_t>::max());
  } else if ((c.dtype == SQLRS_DT_INT32 || c.dtype == SQLRS_DT_INT64 || c.dtype == SQLRS_DT_BOOL) && to == SQLRS_DT_FLOAT64) {
    for (int64_t k = 0; k < c.n; k++) out->f[k] = (double)c.i[k];
  } else if (c.dtype == SQLRS_DT_FLOAT64 && (to == SQLRS_DT_INT64 || to == SQLRS_DT_INT32)) {
    double lo = to == SQLRS_DT_INT64 ? -9223372036854775808.0 : -2147483648.0;
    double hi = to == SQLRS_DT_INT64 ? 9223372036854775808.0 : 2147483648.0;
    for (int64_t k = 0; k < c.n; k++) {
      double v = std::trunc(c.f[k]);
      if (!(v >= lo && v < hi)) {  // also catches NaN
        out->set_null(k);
        out->i[k] = 0;
      } else {
        out->i[k] = (int64_t)v;
      }
    }
  } else if ((c.dtype == SQLRS_DT_INT32 || c.dtype == SQLRS_DT_INT64) && to == SQLRS_DT_BOOL) {
    for (int64_t k = 0; k < c.n; k++) out->i[k] = c.i[k] != 0;
  } else if (c.dtype == SQLRS_DT_FLOAT64 && to == SQLRS_DT_BOOL) {
    for (int64_t k = 0; k < c.n; k++) out->i[k] = c.f[k] != 0.0;
  } else {
    fail(SQLRS_ERR_ARROW, std::string("Casting from ") + dtype_name(c.dtype) + " to " + dtype_name(to) + " not supported");
  }
  out->normalize();
  return out;
}

// BoundExpr::eval_column, evaluator.rs:13-28, over the flattened (postfix) tree
inline ColPtr eval_expr(const Expr& e, const Batch& batch) {
  if (e.empty()) fail(SQLRS_ERR_INVALID_ARG, "empty expression");
  std::vector<ColPtr> stack;
  for (const ExprNode& n : e) {
    switch (n.op) {
      case SQLRS_OP_INPUT_REF:
        if (n.index < 0 || n.index >= (int)batch.cols.size())
          fail(SQLRS_ERR_INTERNAL, "InputRef index out of bounds");
        stack.push_back(batch.cols[n.index]);
        break;
      case SQLRS_OP_CONSTANT:
        stack.push_back(scalar_to_column(constant_scalar(n), batch.n));
        break;
      case SQLRS_OP_CAST: {
        if (stack.empty()) fail(SQLRS_ERR_INVALID_ARG, "malformed expression");
        ColPtr a = stack.back();
        stack.pop_back();
        stack.push_back(cast_column(a, n.dtype));
        break;
      }
      default: {
        if (stack.size() < 2) fail(SQLRS_ERR_INVALID_ARG, "malformed expression");
        ColPtr r = stack.back();
        stack.pop_back();
        ColPtr l = stack.back();
        stack.pop_back();
        if (n.op >= SQLRS_OP_ADD && n.op <= SQLRS_OP_DIV_CHECKED) stack.push_back(arithmetic(*l, *r, n.op));
        else if (n.op >= SQLRS_OP_GT && n.op <= SQLRS_OP_NE) stack.push_back(comparison(*l, *r, n.op));
        else if (n.op == SQLRS_OP_AND || n.op == SQLRS_OP_OR) stack.push_back(kleene(*l, *r, n.op));
        else fail(SQLRS_ERR_UNSUPPORTED, "todo!: unsupported binary operator");  // array_compute.rs:88
      }
    }
  }
  if (stack.size() != 1) fail(SQLRS_ERR_INVALID_ARG, "malformed expression");
  return stack.back();
}

// static result type of an expression (what eval_field reports, evaluator.rs:30-64)
inline int expr_dtype(const Expr& e, const Batch& batch) {
  const ExprNode& n = e.back();
  if (n.op == SQLRS_OP_INPUT_REF) return batch.fields.at(n.index).dtype;
  return n.dtype;
}

}  // namespace oracle
