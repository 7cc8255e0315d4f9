// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the sqlrs v1 executor hot path.
// Nothing under sqlrs_b200/ may include, link or call this; only tests/, the smoke check
// and bench.py's cpu_baseline / --impl reference legs do.
//
// Column / batch / scalar model: the type universe of src/types/mod.rs:23-36
// (Null, Boolean, Float64, Int32, Int64, String) over Arrow-like columns
// (values + optional validity), immutable and shared like `ArrayRef = Arc<dyn Array>`.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/sqlrs_b200.h"

namespace oracle {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
[[noreturn]] inline void fail(int code, const std::string& m) { throw Error(code, m); }

inline const char* dtype_name(int dt) {
  switch (dt) {
    case SQLRS_DT_NULL: return "Null";
    case SQLRS_DT_BOOL: return "Boolean";
    case SQLRS_DT_INT32: return "Int32";
    case SQLRS_DT_INT64: return "Int64";
    case SQLRS_DT_FLOAT64: return "Float64";
    case SQLRS_DT_UTF8: return "Utf8";
  }
  return "?";
}

// One column.  Int32 values are kept sign-extended in `i` (arithmetic re-wraps to 32 bits),
// Boolean as 0/1 in `i`.  `valid` empty <=> null_count == 0 (Arrow: validity buffer absent).
struct Column {
  int dtype = SQLRS_DT_NULL;
  int64_t n = 0;
  std::vector<uint8_t> valid;
  std::vector<int64_t> i;
  std::vector<double> f;
  std::vector<std::string> s;

  bool is_valid(int64_t r) const {
    if (dtype == SQLRS_DT_NULL) return false;
    return valid.empty() || valid[r];
  }
  int64_t null_count() const {
    if (dtype == SQLRS_DT_NULL) return n;
    if (valid.empty()) return 0;
    int64_t c = 0;
    for (uint8_t v : valid) c += (v == 0);
    return c;
  }
  void alloc(int dt, int64_t rows) {
    dtype = dt;
    n = rows;
    if (dt == SQLRS_DT_FLOAT64) f.assign(rows, 0.0);
    else if (dt == SQLRS_DT_UTF8) s.assign(rows, std::string());
    else if (dt != SQLRS_DT_NULL) i.assign(rows, 0);
  }
  void set_null(int64_t r) {
    if (valid.empty()) valid.assign(n, 1);
    valid[r] = 0;
  }
  // drop the validity vector again if nothing is null (keeps null_count()==0 <=> valid.empty())
  void normalize() {
    if (!valid.empty() && null_count() == 0) valid.clear();
  }
};
using ColPtr = std::shared_ptr<const Column>;

struct Field {
  std::string name;
  int dtype = SQLRS_DT_NULL;
  bool nullable = true;
};

// RecordBatch
struct Batch {
  std::vector<Field> fields;
  std::vector<ColPtr> cols;
  int64_t n = 0;
};

// ScalarValue (src/types/mod.rs:23-36): typed, nullable single value.
struct Scalar {
  int dtype = SQLRS_DT_NULL;
  bool is_null = true;
  int64_t i = 0;
  double f = 0;
  std::string s;

  static Scalar null_of(int dt) {  // ScalarValue::from(&DataType)  types/mod.rs:50-60
    Scalar v;
    v.dtype = dt;
    return v;
  }
  // ScalarValue::try_from_array  types/mod.rs:63-78
  static Scalar from_column(const Column& c, int64_t r) {
    Scalar v = null_of(c.dtype);
    if (!c.is_valid(r)) return v;
    v.is_null = false;
    if (c.dtype == SQLRS_DT_FLOAT64) v.f = c.f[r];
    else if (c.dtype == SQLRS_DT_UTF8) v.s = c.s[r];
    else v.i = c.i[r];
    return v;
  }
  // PartialEq / Hash (types/mod.rs:170-212): floats compare via OrderedFloat (NaN == NaN, -0.0 == 0.0)
  bool equals(const Scalar& o) const {
    if (dtype != o.dtype) return false;
    if (dtype == SQLRS_DT_NULL) return true;
    if (is_null != o.is_null) return false;
    if (is_null) return true;
    if (dtype == SQLRS_DT_FLOAT64) return (f == o.f) || (f != f && o.f != o.f);
    if (dtype == SQLRS_DT_UTF8) return s == o.s;
    return i == o.i;
  }
};

// append a scalar to a column under construction (append_scalar_value_for_builder, types/mod.rs:236-273)
inline void append_scalar(Column& c, const Scalar& v) {
  if (v.dtype == SQLRS_DT_NULL)
    fail(SQLRS_ERR_ARROW, "NotYetImplemented: not support Null as group by key");
  if (c.dtype != v.dtype) fail(SQLRS_ERR_INTERNAL, "builder/scalar type mismatch");
  int64_t r = c.n++;
  if (c.dtype == SQLRS_DT_FLOAT64) c.f.push_back(v.f);
  else if (c.dtype == SQLRS_DT_UTF8) c.s.push_back(v.s);
  else c.i.push_back(v.i);
  if (!c.valid.empty()) c.valid.push_back(v.is_null ? 0 : 1);
  else if (v.is_null) {
    c.valid.assign(c.n, 1);
    c.valid[r] = 0;
  }
}

// build_scalar_value_array (types/mod.rs:214-223): a constant materialised to `n` rows.
inline ColPtr scalar_to_column(const Scalar& v, int64_t n) {
  auto c = std::make_shared<Column>();
  c->alloc(v.dtype, n);
  if (v.dtype == SQLRS_DT_NULL) return c;
  if (v.is_null) {
    c->valid.assign(n, 0);
    return c;
  }
  if (v.dtype == SQLRS_DT_FLOAT64) std::fill(c->f.begin(), c->f.end(), v.f);
  else if (v.dtype == SQLRS_DT_UTF8) std::fill(c->s.begin(), c->s.end(), v.s);
  else std::fill(c->i.begin(), c->i.end(), v.i);
  return c;
}

// arrow compute::take with a nullable index array (index < 0 encodes NULL -> NULL row)
template <typename Idx>
inline ColPtr take(const Column& src, const std::vector<Idx>& idx, const std::vector<uint8_t>* idx_valid = nullptr) {
  auto out = std::make_shared<Column>();
  int64_t m = (int64_t)idx.size();
  out->alloc(src.dtype, m);
  if (src.dtype == SQLRS_DT_NULL) return out;
  bool any_null = !src.valid.empty() || idx_valid != nullptr;
  if (any_null) out->valid.assign(m, 1);
  for (int64_t k = 0; k < m; k++) {
    if (idx_valid && !(*idx_valid)[k]) {
      out->valid[k] = 0;
      continue;
    }
    int64_t r = (int64_t)idx[k];
    if (!src.valid.empty() && !src.valid[r]) out->valid[k] = 0;
    if (src.dtype == SQLRS_DT_FLOAT64) out->f[k] = src.f[r];
    else if (src.dtype == SQLRS_DT_UTF8) out->s[k] = src.s[r];
    else out->i[k] = src.i[r];
  }
  out->normalize();
  return out;
}

// arrow compute::filter / filter_record_batch: keep rows whose mask is valid && true
inline ColPtr filter_column(const Column& src, const std::vector<uint8_t>& keep, int64_t n_keep) {
  auto out = std::make_shared<Column>();
  out->alloc(src.dtype, n_keep);
  if (src.dtype == SQLRS_DT_NULL) return out;
  if (!src.valid.empty()) out->valid.assign(n_keep, 1);
  int64_t k = 0;
  for (int64_t r = 0; r < src.n; r++) {
    if (!keep[r]) continue;
    if (!src.valid.empty() && !src.valid[r]) out->valid[k] = 0;
    if (src.dtype == SQLRS_DT_FLOAT64) out->f[k] = src.f[r];
    else if (src.dtype == SQLRS_DT_UTF8) out->s[k] = src.s[r];
    else out->i[k] = src.i[r];
    k++;
  }
  out->normalize();
  return out;
}

// compute::concat_batches on one column position
inline ColPtr concat_columns(const std::vector<ColPtr>& parts) {
  auto out = std::make_shared<Column>();
  if (parts.empty()) return out;
  int dt = parts[0]->dtype;
  int64_t total = 0;
  bool any_null = false;
  for (auto& p : parts) {
    if (p->dtype != dt) fail(SQLRS_ERR_ARROW, "concat_batches: column type mismatch");
    total += p->n;
    any_null |= !p->valid.empty();
  }
  out->dtype = dt;
  out->n = total;
  if (dt == SQLRS_DT_NULL) return out;
  if (any_null) out->valid.reserve(total);
  for (auto& p : parts) {
    if (dt == SQLRS_DT_FLOAT64) out->f.insert(out->f.end(), p->f.begin(), p->f.end());
    else if (dt == SQLRS_DT_UTF8) out->s.insert(out->s.end(), p->s.begin(), p->s.end());
    else out->i.insert(out->i.end(), p->i.begin(), p->i.end());
    if (any_null) {
      if (p->valid.empty()) out->valid.insert(out->valid.end(), p->n, 1);
      else out->valid.insert(out->valid.end(), p->valid.begin(), p->valid.end());
    }
  }
  return out;
}

inline ColPtr null_column(int dtype, int64_t n) {  // arrow new_null_array
  auto c = std::make_shared<Column>();
  c->alloc(dtype, n);
  if (dtype != SQLRS_DT_NULL) c->valid.assign(n, 0);
  return c;
}

}  // namespace oracle
