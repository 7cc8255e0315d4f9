// sqlrs_b200 — expression bytecode -> CUDA source of a straight-line "row program".
//
// The reference evaluates a BoundExpr tree one arrow kernel + one full-length temporary per
// node (src/executor/evaluator.rs:13-28, array_compute.rs:70-90).  Here every expression an
// operator needs (predicate, group keys, aggregate arguments, join keys) is flattened into ONE
// typed, common-subexpression-eliminated sequence of C statements over the values of a single
// row; the hand-written kernel skeletons in csrc/jit/*.cuh call it once per row, so a whole
// Filter->Agg pipeline reads every input column exactly once and never materialises a
// temporary.  Type checking happens here and mirrors the reference's failure modes
// (status codes of include/sqlrs_b200.h).
#pragma once
#include <map>
#include <sstream>

#include "common.hpp"

namespace sq {

struct ColInfo {
  int dtype = SQLRS_DT_NULL;
  bool has_valid = false;  // a validity bitmap is present (null_count != 0)
  bool nullable = true;    // the schema field is declared nullable (stable across batches)
};

// a compiled value: C identifiers v<id> (typed) and n<id> (bool "is valid")
struct Val {
  int id = -1;
  int dtype = SQLRS_DT_NULL;
  bool maybe_null = false;   // false: statically never NULL in THIS batch (no validity bitmap upstream)
  bool decl_null = false;    // may be NULL in some batch of this schema (drives state layouts, which must not change between batches)
  bool always_null = false;  // typed NULL literal / Null column
};

struct ExprNodeCopy {
  int op = 0, dtype = 0, index = 0, is_null = 0;
  int64_t imm_bits = 0;
  bool has_str = false;
};
using ExprCopy = std::vector<ExprNodeCopy>;
ExprCopy copy_expr(const sqlrs_expr* e);

const char* ctype_of(int dtype);  // C type used for a value of that dtype in generated code

class RowProgram {
 public:
  explicit RowProgram(std::vector<ColInfo> cols) : cols_(std::move(cols)) {}
  // joined mode (fused probe -> consumer pipelines): InputRef k addresses the joined row of a hash join —
  // k < build_cols.size() loads build-side column k at the matched build row `b` (SQ_LDB_* macros), the rest
  // loads probe-side column k - build_cols.size() at the scan row `r`
  RowProgram(std::vector<ColInfo> build_cols, std::vector<ColInfo> probe_cols)
      : cols_(std::move(probe_cols)), build_cols_(std::move(build_cols)), joined_(true) {}
  // compiles one expression; statements are appended to the body.  `err_class` selects the
  // error flag a runtime failure (divide by zero) of this expression raises: 0 = applies to every
  // row, 1 = only to rows that pass the predicate (expressions above a fused Filter).
  Val compile(const ExprCopy& e, int err_class);
  // create_hashes (hash_utils.rs:161-220) over already compiled key values -> u64 value id
  int emit_row_hash(const std::vector<Val>& keys);
  // raw 64-bit pattern of a value (i32 sign-extended, bool 0/1, f64 bits) -> u64 value id
  int emit_raw_bits(const Val& v);
  // cheap placement hash over raw key bits + null flags (only valid where key tuples are compared too)
  int emit_mix_hash(const std::vector<int>& raw_ids, const std::vector<Val>& keys);
  // the same hash as C statements over an array `kb[]` of raw key bits of NON-NULL keys of these dtypes (the fused probe
  // kernels queue the key bits of a candidate row and re-derive its hash instead of re-reading the row)
  static std::string mix_hash_of_bits_source(const std::vector<int>& dtypes);
  Val cast(const Val& a, int to);  // arrow compute::cast (also used by SUM: sum.rs:54)
  int fresh() { return next_id_++; }
  std::ostringstream& body() { return body_; }
  std::string body_str() const { return body_.str(); }
  const std::vector<ColInfo>& cols() const { return cols_; }
  bool uses_error_flag(int cls) const { return err_used_[cls]; }
  std::string signature() const;  // schema part of the JIT cache key
  int n_loaded_columns() const {   // distinct input columns the program reads
    int n = 0;
    for (const auto& kv : cse_) n += kv.first.rfind("col", 0) == 0;
    return n;
  }

 private:
  Val load_column(int index);
  Val constant(const ExprNodeCopy& n);
  Val arithmetic(const Val& l, const Val& r, int op, int err_class);
  Val comparison(const Val& l, const Val& r, int op);
  Val kleene(const Val& l, const Val& r, int op);
  Val define(int dtype, const std::string& value_expr, const std::string& valid_expr, bool maybe_null, bool decl_null,
             const std::string& cse_key);

  std::vector<ColInfo> cols_;
  std::vector<ColInfo> build_cols_;
  bool joined_ = false;
  std::ostringstream body_;
  std::map<std::string, Val> cse_;
  int next_id_ = 0;
  bool err_used_[2] = {false, false};
};

// set once a *_checked operator has been compiled: the operators' "row error" flag then also stands for integer overflow
extern bool g_checked_arithmetic_compiled;
inline const char* arithmetic_error_text() {
  return g_checked_arithmetic_compiled ? "Compute error: Overflow happened (checked arithmetic) or Divide by zero error" : "Divide by zero error";
}

// C literal helpers
std::string lit_i64(int64_t v);
std::string lit_f64_bits(int64_t bits);

}  // namespace sq
