// ORACLE — TEST INFRASTRUCTURE ONLY.
// Step-for-step CPU restatement of the sqlrs v1 operators on the hot path:
//   FilterExecutor      src/executor/filter.rs:14-26
//   Accumulators        src/executor/aggregate/{mod.rs:27-49, sum.rs:16-132, count.rs:10-58, min_max.rs:47-157}
//   SimpleAggExecutor   src/executor/aggregate/simple_agg.rs:27-65
//   HashAggExecutor     src/executor/aggregate/hash_agg.rs:33-150
//   HashJoinExecutor    src/executor/join/hash_join.rs:25-127,147-323
// Same per-batch structure as the reference (hash rows -> map -> per-group take -> update_batch;
// build map -> concat -> probe -> gather twice), so that timing it is a fair "sqlrs-equivalent
// CPU path" and so that the quirks K1 (count overwrites), K2 (hash-only identity) and K3
// (NULL key keeps hash 0) fall out of the structure instead of being special-cased.
// The SQL-semantics switches (count_mode / match_mode) are explicit deviations.
#pragma once
#include <algorithm>
#include <unordered_map>

#include "ahash.hpp"
#include "expr.hpp"

namespace oracle {

struct Options {
  int count_mode = SQLRS_COUNT_REFERENCE_OVERWRITE;
  int match_mode = SQLRS_MATCH_HASH_ONLY;
};
inline Options copy_options(const sqlrs_options* o) {
  Options r;
  if (o) {
    r.count_mode = o->count_mode;
    r.match_mode = o->match_mode;
  }
  return r;
}

// ------------------------------------------------------------------ Filter
// filter.rs:16-25: mask = eval; must be Boolean (else the reference panics via expect);
// filter_record_batch keeps rows whose mask is valid and true; row order preserved.
inline Batch filter_batch(const Expr& predicate, const Batch& batch) {
  ColPtr mask = eval_expr(predicate, batch);
  if (mask->dtype != SQLRS_DT_BOOL)
    fail(SQLRS_ERR_INTERNAL, "filter executor expected evaluate boolean array");
  if (mask->n != batch.n) fail(SQLRS_ERR_ARROW, "filter mask length mismatch");
  std::vector<uint8_t> keep(batch.n);
  int64_t n_keep = 0;
  for (int64_t r = 0; r < batch.n; r++) {
    keep[r] = (uint8_t)(mask->is_valid(r) && mask->i[r] != 0);
    n_keep += keep[r];
  }
  Batch out;
  out.fields = batch.fields;
  out.n = n_keep;
  for (const ColPtr& c : batch.cols) out.cols.push_back(filter_column(*c, keep, n_keep));
  return out;
}

// ------------------------------------------------------------------ Accumulators
struct AggDesc {
  int func = 0, distinct = 0, return_dtype = 0;
  Expr arg;
  std::string name;
};
inline AggDesc copy_agg(const sqlrs_agg_desc& d) {
  AggDesc a;
  a.func = d.func;
  a.distinct = d.distinct;
  a.return_dtype = d.return_dtype;
  a.arg = copy_expr(&d.arg);
  a.name = d.name ? d.name : "";
  return a;
}

struct Accumulator {
  int func, distinct, dtype, count_mode;
  Scalar result;                 // Sum / Min / Max state
  int64_t count = 0;             // CountAccumulator::result
  std::vector<Scalar> distinct_values;  // HashSet<ScalarValue> (order irrelevant for ints)

  Accumulator(const AggDesc& d, int count_mode_)
      : func(d.func), distinct(d.distinct), dtype(d.return_dtype), count_mode(count_mode_),
        result(Scalar::null_of(d.return_dtype)) {}

  // sum_result, sum.rs:64-85 + typed_sum! :25-33
  static Scalar sum_result(const Scalar& l, const Scalar& r) {
    Scalar out = l;
    bool ok = (l.dtype == SQLRS_DT_FLOAT64 && is_numeric(r.dtype)) ||
              (l.dtype == SQLRS_DT_INT64 && (r.dtype == SQLRS_DT_INT64 || r.dtype == SQLRS_DT_INT32));
    if (!ok) {
      // the reference's Int32 accumulator: ScalarValue::Int32 + Int32 hits `unimplemented!`
      fail(SQLRS_ERR_UNSUPPORTED, std::string("not expected ") + dtype_name(l.dtype) + " and " +
                                      dtype_name(r.dtype) + " for sum");
    }
    if (r.is_null) return out;
    double rf = r.dtype == SQLRS_DT_FLOAT64 ? r.f : (double)r.i;
    if (l.is_null) {
      out.is_null = false;
      if (l.dtype == SQLRS_DT_FLOAT64) out.f = rf;
      else out.i = r.i;
      return out;
    }
    if (l.dtype == SQLRS_DT_FLOAT64) out.f = l.f + rf;
    else out.i = (int64_t)((uint64_t)l.i + (uint64_t)r.i);  // release build: wraps
    return out;
  }

  // arrow compute::sum: skips NULLs, None if there is no valid value; ints wrap.
  static Scalar sum_batch(const Column& c) {
    Scalar s = Scalar::null_of(c.dtype);
    if (!is_numeric(c.dtype)) fail(SQLRS_ERR_UNSUPPORTED, std::string("unsupported sum type: ") + dtype_name(c.dtype));
    bool any = false;
    if (c.dtype == SQLRS_DT_FLOAT64) {
      double acc = 0;
      for (int64_t r = 0; r < c.n; r++)
        if (c.is_valid(r)) {
          acc += c.f[r];
          any = true;
        }
      s.f = acc;
    } else {
      uint64_t acc = 0;
      for (int64_t r = 0; r < c.n; r++)
        if (c.is_valid(r)) {
          acc += (uint64_t)c.i[r];
          any = true;
        }
      s.i = wrap_int(c.dtype, (int64_t)acc);
    }
    s.is_null = !any;
    return s;
  }

  static Scalar min_max_batch(const Column& c, bool is_min) {
    Scalar s = Scalar::null_of(c.dtype);
    if (!(is_numeric(c.dtype) || c.dtype == SQLRS_DT_UTF8))
      fail(SQLRS_ERR_UNSUPPORTED, std::string("unsupported min/max type: ") + dtype_name(c.dtype));
    for (int64_t r = 0; r < c.n; r++) {
      if (!c.is_valid(r)) continue;
      if (s.is_null) {
        s = Scalar::from_column(c, r);
        continue;
      }
      if (c.dtype == SQLRS_DT_FLOAT64) {
        // arrow min/max on floats: NaN-aware total order is UNPINNED; IEEE compare, NaN ignored
        if (is_min ? (c.f[r] < s.f) : (c.f[r] > s.f)) s.f = c.f[r];
      } else if (c.dtype == SQLRS_DT_UTF8) {
        if (is_min ? (c.s[r] < s.s) : (c.s[r] > s.s)) s.s = c.s[r];
      } else {
        if (is_min ? (c.i[r] < s.i) : (c.i[r] > s.i)) s.i = c.i[r];
      }
    }
    return s;
  }

  // min_max!, min_max.rs:91-109
  static Scalar min_max_merge(const Scalar& v, const Scalar& d, bool is_min) {
    if (v.dtype != d.dtype || v.dtype == SQLRS_DT_BOOL || v.dtype == SQLRS_DT_NULL)
      fail(SQLRS_ERR_UNSUPPORTED, std::string("unsupported min_max scalar type: ") + dtype_name(v.dtype));
    if (d.is_null) return v;
    if (v.is_null) return d;
    Scalar out = v;
    if (v.dtype == SQLRS_DT_FLOAT64) out.f = is_min ? std::fmin(v.f, d.f) : std::fmax(v.f, d.f);  // f64::min/max
    else if (v.dtype == SQLRS_DT_UTF8) out.s = is_min ? std::min(v.s, d.s) : std::max(v.s, d.s);
    else out.i = is_min ? std::min(v.i, d.i) : std::max(v.i, d.i);
    return out;
  }

  void update_batch(const ColPtr& array) {
    const Column& c = *array;
    if (distinct && (func == SQLRS_AGG_COUNT || func == SQLRS_AGG_SUM)) {
      // count.rs:44-53 / sum.rs:114-123: every row (NULLs included) goes into the set
      for (int64_t r = 0; r < c.n; r++) {
        Scalar v = Scalar::from_column(c, r);
        bool seen = false;
        for (const Scalar& o : distinct_values)
          if (o.equals(v)) {
            seen = true;
            break;
          }
        if (!seen) distinct_values.push_back(v);
      }
      return;
    }
    switch (func) {
      case SQLRS_AGG_COUNT: {
        int64_t cnt = c.n - c.null_count();
        if (count_mode == SQLRS_COUNT_REFERENCE_OVERWRITE) count = cnt;  // count.rs:22 (quirk K1)
        else count += cnt;
        break;
      }
      case SQLRS_AGG_SUM: {
        ColPtr casted = cast_column(array, dtype);  // sum.rs:54
        result = sum_result(result, sum_batch(*casted));
        break;
      }
      case SQLRS_AGG_MIN: result = min_max_merge(result, min_max_batch(c, true), true); break;
      case SQLRS_AGG_MAX: result = min_max_merge(result, min_max_batch(c, false), false); break;
      default: fail(SQLRS_ERR_INVALID_ARG, "unknown aggregate function");
    }
  }

  Scalar evaluate() const {
    if (func == SQLRS_AGG_COUNT) {
      Scalar s = Scalar::null_of(SQLRS_DT_INT64);
      s.is_null = false;
      s.i = distinct ? (int64_t)distinct_values.size() : count;
      return s;
    }
    if (func == SQLRS_AGG_SUM && distinct) {  // sum.rs:125-131
      Scalar sum = Scalar::null_of(dtype);
      for (const Scalar& v : distinct_values) sum = sum_result(sum, v);
      return sum;
    }
    return result;
  }
};

inline int agg_output_dtype(const AggDesc& d) { return d.func == SQLRS_AGG_COUNT ? SQLRS_DT_INT64 : d.return_dtype; }

// ------------------------------------------------------------------ SimpleAgg
struct SimpleAgg {
  std::vector<AggDesc> aggs;
  Options opt;
  std::vector<Accumulator> accs;
  bool seen_batch = false;

  SimpleAgg(std::vector<AggDesc> a, Options o) : aggs(std::move(a)), opt(o) {
    for (const AggDesc& d : aggs) accs.emplace_back(d, opt.count_mode);
  }
  void push(const Batch& batch) {  // simple_agg.rs:34-54
    std::vector<ColPtr> columns;
    for (const AggDesc& d : aggs) columns.push_back(eval_expr(d.arg, batch));
    seen_batch = true;
    for (size_t k = 0; k < accs.size(); k++) accs[k].update_batch(columns[k]);
  }
  Batch finish() {  // :56-64
    if (!seen_batch) fail(SQLRS_ERR_INTERNAL, "called `Option::unwrap()` on a `None` value (no input batch)");
    Batch out;
    out.n = 1;
    for (size_t k = 0; k < accs.size(); k++) {
      Scalar v = accs[k].evaluate();
      out.cols.push_back(scalar_to_column(v, 1));
      out.fields.push_back(Field{aggs[k].name, v.dtype, true});
    }
    return out;
  }
};

// ------------------------------------------------------------------ HashAgg
struct HashAgg {
  std::vector<AggDesc> aggs;
  std::vector<Expr> group_by;
  std::vector<std::string> group_names;
  Options opt;

  struct Group {
    uint64_t hash;
    std::vector<Scalar> keys;
    std::vector<Accumulator> accs;
    int64_t first_row = 0;  // global row id of the group's first row (orders merged partial groups)
  };
  int64_t rows_seen = 0;
  std::vector<Group> groups;                               // group_hashs order (first appearance)
  std::unordered_map<uint64_t, std::vector<int>> by_hash;  // hash -> group ids (1 id in HASH_ONLY mode)
  std::vector<int> key_dtypes;
  bool seen_batch = false;

  HashAgg(std::vector<AggDesc> a, std::vector<Expr> g, std::vector<std::string> names, Options o)
      : aggs(std::move(a)), group_by(std::move(g)), group_names(std::move(names)), opt(o) {}

  int find_or_create(uint64_t hash, const std::vector<ColPtr>& keys, int64_t row) {
    auto& ids = by_hash[hash];
    if (opt.match_mode == SQLRS_MATCH_HASH_ONLY) {
      if (!ids.empty()) return ids[0];  // hash_agg.rs:87: identity is the hash alone (quirk K2)
    } else {
      for (int id : ids) {
        bool eq = true;
        for (size_t k = 0; k < keys.size() && eq; k++) eq = groups[id].keys[k].equals(Scalar::from_column(*keys[k], row));
        if (eq) return id;
      }
    }
    Group g;
    g.hash = hash;
    g.first_row = rows_seen + row;
    for (const ColPtr& k : keys) g.keys.push_back(Scalar::from_column(*k, row));  // :92-96
    for (const AggDesc& d : aggs) g.accs.emplace_back(d, opt.count_mode);         // :89
    groups.push_back(std::move(g));
    ids.push_back((int)groups.size() - 1);
    return (int)groups.size() - 1;
  }

  void push(const Batch& batch) {
    std::vector<ColPtr> columns, keys;
    for (const AggDesc& d : aggs) columns.push_back(eval_expr(d.arg, batch));  // :63-66
    for (const Expr& e : group_by) keys.push_back(eval_expr(e, batch));        // :69-73
    if (!seen_batch) {
      for (const ColPtr& k : keys) key_dtypes.push_back(k->dtype);
      seen_batch = true;
    }
    std::vector<uint64_t> hashes(batch.n, 0);  // :76
    create_hashes(keys, hashes);
    // :86-110 — per-batch row lists per group, in first-touch order (the reference iterates a
    // std HashMap here; the order is result-irrelevant because groups are independent)
    std::vector<int> touched;
    std::unordered_map<int, std::vector<uint32_t>> rows_of;
    for (int64_t row = 0; row < batch.n; row++) {
      int g = find_or_create(hashes[row], keys, row);
      auto it = rows_of.find(g);
      if (it == rows_of.end()) {
        touched.push_back(g);
        it = rows_of.emplace(g, std::vector<uint32_t>()).first;
      }
      it->second.push_back((uint32_t)row);
    }
    for (int g : touched) {  // :113-121
      const std::vector<uint32_t>& idx = rows_of[g];
      for (size_t k = 0; k < aggs.size(); k++) groups[g].accs[k].update_batch(take(*columns[k], idx));
    }
    rows_seen += batch.n;
  }

  // ---- partial / final split (multi-process group-by; NOT part of the reference):
  // table 0: un-finalised groups as a batch [hash, min_row, keys..., per aggregate: state value, count];
  // table 1 + j (the j-th DISTINCT COUNT / SUM): the elements of the per-group sets, [hash, min_row, keys..., element]
  static bool is_distinct_agg(const AggDesc& d) { return d.distinct && (d.func == SQLRS_AGG_COUNT || d.func == SQLRS_AGG_SUM); }
  std::vector<size_t> distinct_aggs() const {
    std::vector<size_t> out;
    for (size_t k = 0; k < aggs.size(); k++)
      if (is_distinct_agg(aggs[k])) out.push_back(k);
    return out;
  }
  int partial_tables() const { return 1 + (int)distinct_aggs().size(); }
  static Scalar i64_scalar(int64_t v) {
    Scalar s = Scalar::null_of(SQLRS_DT_INT64);
    s.is_null = false;
    s.i = v;
    return s;
  }

  Batch export_partials(int table = 0) const {
    if (!seen_batch) fail(SQLRS_ERR_INTERNAL, "called `Option::unwrap()` on a `None` value (no input batch)");
    if (opt.count_mode != SQLRS_COUNT_SQL_ACCUMULATE) fail(SQLRS_ERR_UNSUPPORTED, "partial/final COUNT needs SQLRS_COUNT_SQL_ACCUMULATE");
    Batch out;
    std::vector<std::shared_ptr<Column>> cols;
    auto add = [&](const std::string& name, int dtype) {
      auto c = std::make_shared<Column>();
      c->dtype = dtype;
      cols.push_back(c);
      out.fields.push_back(Field{name, dtype, true});
    };
    add("hash", SQLRS_DT_INT64);
    add("min_row", SQLRS_DT_INT64);
    for (size_t k = 0; k < group_by.size(); k++) add("key" + std::to_string(k), key_dtypes[k]);
    int64_t n = 0;
    if (table == 0) {
      for (size_t k = 0; k < aggs.size(); k++) {
        add("state" + std::to_string(k), agg_output_dtype(aggs[k]));
        add("count" + std::to_string(k), SQLRS_DT_INT64);
      }
      for (const Group& g : groups) {
        size_t c = 0;
        append_scalar(*cols[c++], i64_scalar((int64_t)g.hash));
        append_scalar(*cols[c++], i64_scalar(g.first_row));
        for (const Scalar& k : g.keys) append_scalar(*cols[c++], k);
        for (size_t k = 0; k < g.accs.size(); k++) {
          const Accumulator& a = g.accs[k];
          // a DISTINCT accumulator's state is its set: table 1 + j carries it
          append_scalar(*cols[c++], is_distinct_agg(aggs[k]) ? Scalar::null_of(agg_output_dtype(aggs[k])) : a.evaluate());
          append_scalar(*cols[c++], i64_scalar(is_distinct_agg(aggs[k]) ? 0 : a.count));
        }
        n++;
      }
    } else {
      const std::vector<size_t> da = distinct_aggs();
      if (table < 0 || (size_t)table > da.size()) fail(SQLRS_ERR_INVALID_ARG, "partials table index out of range");
      const size_t k = da[(size_t)table - 1];
      int elem_dtype = SQLRS_DT_INT64;  // (no element anywhere: an empty batch of any type)
      for (const Group& g : groups)
        if (!g.accs[k].distinct_values.empty()) elem_dtype = g.accs[k].distinct_values[0].dtype;
      add("element", elem_dtype);
      for (const Group& g : groups)
        for (const Scalar& v : g.accs[k].distinct_values) {
          size_t c = 0;
          append_scalar(*cols[c++], i64_scalar((int64_t)g.hash));
          append_scalar(*cols[c++], i64_scalar(g.first_row));
          for (const Scalar& key : g.keys) append_scalar(*cols[c++], key);
          append_scalar(*cols[c++], v);
          n++;
        }
    }
    out.n = n;
    for (auto& c : cols) {
      c->normalize();
      out.cols.push_back(c);
    }
    return out;
  }
  void clear_partials(int table = 0) {
    if (table == 0) {
      groups.clear();
      by_hash.clear();
      return;
    }
    const std::vector<size_t> da = distinct_aggs();
    if (table < 0 || (size_t)table > da.size()) fail(SQLRS_ERR_INVALID_ARG, "partials table index out of range");
    for (Group& g : groups) g.accs[da[(size_t)table - 1]].distinct_values.clear();
  }
  // the group of partial row r (created if this process has not seen it), first-appearance bookkeeping included
  Group& merge_group(const Batch& p, const std::vector<ColPtr>& keys, int64_t r) {
    const size_t K = group_by.size();
    const int64_t saved = rows_seen;
    rows_seen = p.cols[1]->i[r];  // find_or_create stamps first_row = rows_seen + row
    const size_t before = groups.size();
    int gid = find_or_create((uint64_t)p.cols[0]->i[r], keys, r);
    if (groups.size() != before) groups[gid].first_row = p.cols[1]->i[r];
    rows_seen = saved;
    Group& g = groups[gid];
    if (p.cols[1]->i[r] < g.first_row) {
      g.first_row = p.cols[1]->i[r];
      if (opt.match_mode == SQLRS_MATCH_HASH_ONLY)
        for (size_t k = 0; k < K; k++) g.keys[k] = Scalar::from_column(*keys[k], r);
    }
    return g;
  }
  void merge_partials(const Batch& p, int table = 0) {
    const size_t K = group_by.size();
    const std::vector<size_t> da = distinct_aggs();
    if (table < 0 || (size_t)table > da.size()) fail(SQLRS_ERR_INVALID_ARG, "partials table index out of range");
    if (p.cols.size() != (table == 0 ? 2 + K + 2 * aggs.size() : 2 + K + 1)) fail(SQLRS_ERR_INVALID_ARG, "partials batch has the wrong number of columns");
    if (!seen_batch) {
      for (size_t k = 0; k < K; k++) key_dtypes.push_back(p.cols[2 + k]->dtype);
      seen_batch = true;
    }
    std::vector<ColPtr> keys(p.cols.begin() + 2, p.cols.begin() + 2 + K);
    for (int64_t r = 0; r < p.n; r++) {
      Group& g = merge_group(p, keys, r);
      if (table > 0) {  // one set element: insert unless present (count.rs:44-53)
        Accumulator& a = g.accs[da[(size_t)table - 1]];
        Scalar v = Scalar::from_column(*p.cols[2 + K], r);
        bool seen = false;
        for (const Scalar& o : a.distinct_values) seen = seen || o.equals(v);
        if (!seen) a.distinct_values.push_back(v);
        continue;
      }
      for (size_t k = 0; k < aggs.size(); k++) {
        if (is_distinct_agg(aggs[k])) continue;
        Accumulator& a = g.accs[k];
        Scalar v = Scalar::from_column(*p.cols[2 + K + 2 * k], r);
        a.count += p.cols[2 + K + 2 * k + 1]->i[r];
        if (a.func == SQLRS_AGG_SUM) a.result = Accumulator::sum_result(a.result, v);
        else if (a.func == SQLRS_AGG_MIN) a.result = Accumulator::min_max_merge(a.result, v, true);
        else if (a.func == SQLRS_AGG_MAX) a.result = Accumulator::min_max_merge(a.result, v, false);
      }
    }
  }

  Batch finish() {  // :124-149
    if (!seen_batch) fail(SQLRS_ERR_INTERNAL, "called `Option::unwrap()` on a `None` value (no input batch)");
    Batch out;
    std::vector<std::shared_ptr<Column>> cols;
    for (size_t k = 0; k < group_by.size(); k++) {
      auto c = std::make_shared<Column>();
      c->dtype = key_dtypes[k];
      cols.push_back(c);
      out.fields.push_back(Field{k < group_names.size() ? group_names[k] : "", key_dtypes[k], true});
    }
    for (const AggDesc& d : aggs) {
      auto c = std::make_shared<Column>();
      c->dtype = agg_output_dtype(d);
      cols.push_back(c);
      out.fields.push_back(Field{d.name, c->dtype, true});
    }
    // first-appearance order; already sorted unless partial groups were merged in
    std::vector<const Group*> ordered;
    for (const Group& g : groups) ordered.push_back(&g);
    std::stable_sort(ordered.begin(), ordered.end(), [](const Group* a, const Group* b) { return a->first_row < b->first_row; });
    for (const Group* gp : ordered) {
      const Group& g = *gp;
      for (size_t k = 0; k < g.keys.size(); k++) append_scalar(*cols[k], g.keys[k]);
      for (size_t k = 0; k < g.accs.size(); k++) append_scalar(*cols[g.keys.size() + k], g.accs[k].evaluate());
    }
    out.n = (int64_t)groups.size();
    for (auto& c : cols) {
      c->normalize();
      out.cols.push_back(c);
    }
    return out;
  }
};

// ------------------------------------------------------------------ HashJoin
struct HashJoin {
  int join_type;
  std::vector<Expr> left_keys, right_keys;
  Expr filter;  // empty = None
  std::vector<Field> out_fields;
  Options opt;

  std::unordered_map<uint64_t, std::vector<int64_t>> left_hashmap;  // :156
  std::vector<Batch> left_batches;
  std::vector<std::vector<ColPtr>> left_key_parts;  // evaluated build keys (HASH_AND_KEY only)
  int64_t left_row_offset = 0;
  bool sealed = false;
  Batch left_single;
  std::vector<ColPtr> left_key_cols;
  std::vector<uint8_t> visited_left;

  void build_push(const Batch& batch) {  // :161-181
    if (sealed) fail(SQLRS_ERR_INVALID_ARG, "hash_join: build_push after probe");
    std::vector<ColPtr> keys;
    for (const Expr& e : left_keys) keys.push_back(eval_expr(e, batch));
    std::vector<uint64_t> hashes(batch.n, 0);
    create_hashes(keys, hashes);
    for (int64_t row = 0; row < batch.n; row++) left_hashmap[hashes[row]].push_back(row + left_row_offset);
    left_row_offset += batch.n;
    left_batches.push_back(batch);
    if (opt.match_mode == SQLRS_MATCH_HASH_AND_KEY) left_key_parts.push_back(keys);
  }

  void seal() {
    if (sealed) return;
    sealed = true;
    if (left_batches.empty()) return;
    left_single.fields = left_batches[0].fields;  // concat_batches(&left_batches[0].schema(), ..) :187
    left_single.n = left_row_offset;
    for (size_t c = 0; c < left_single.fields.size(); c++) {
      std::vector<ColPtr> parts;
      for (const Batch& b : left_batches) {
        if (b.cols.size() != left_single.fields.size()) fail(SQLRS_ERR_ARROW, "concat_batches: schema mismatch");
        parts.push_back(b.cols[c]);
      }
      left_single.cols.push_back(concat_columns(parts));
    }
    if (opt.match_mode == SQLRS_MATCH_HASH_AND_KEY)
      for (size_t k = 0; k < left_keys.size(); k++) {
        std::vector<ColPtr> parts;
        for (auto& p : left_key_parts) parts.push_back(p[k]);
        left_key_cols.push_back(concat_columns(parts));
      }
    if (join_type == SQLRS_JOIN_LEFT || join_type == SQLRS_JOIN_FULL) visited_left.assign(left_single.n, 0);
  }

  bool keys_equal(const std::vector<ColPtr>& rk, int64_t rrow, int64_t lrow) const {
    for (size_t k = 0; k < rk.size(); k++) {
      Scalar a = Scalar::from_column(*left_key_cols[k], lrow), b = Scalar::from_column(*rk[k], rrow);
      if (a.is_null || b.is_null || !a.equals(b)) return false;  // SQL: NULL never joins
    }
    return true;
  }

  // build_batch, :25-45.  idx < 0 encodes a NULL index.
  Batch build_batch(const Batch& right, const std::vector<int64_t>& li, const std::vector<uint8_t>& li_valid,
                    const std::vector<uint32_t>& ri) const {
    Batch out;
    out.fields = out_fields;
    out.n = (int64_t)li.size();
    for (const ColPtr& c : left_single.cols) out.cols.push_back(take(*c, li, &li_valid));
    for (const ColPtr& c : right.cols) out.cols.push_back(take(*c, ri));
    check_schema(out);
    return out;
  }

  // RecordBatch::try_new(schema, data): column count / types / nullability must agree
  void check_schema(const Batch& b) const {
    if (b.cols.size() != out_fields.size())
      fail(SQLRS_ERR_ARROW, "number of columns must match number of fields in schema");
    for (size_t c = 0; c < b.cols.size(); c++) {
      if (b.cols[c]->dtype != out_fields[c].dtype)
        fail(SQLRS_ERR_ARROW, std::string("column types must match schema types, expected ") +
                                  dtype_name(out_fields[c].dtype) + " but found " + dtype_name(b.cols[c]->dtype));
      if (!out_fields[c].nullable && b.cols[c]->null_count() > 0)
        fail(SQLRS_ERR_ARROW, "Column '" + out_fields[c].name + "' is declared as non-nullable but contains null values");
    }
  }

  // one probe batch, :208-292.  Returns false when the reference yields nothing (empty build side).
  bool probe(const Batch& right, Batch* result) {
    seal();
    if (left_batches.empty()) return false;  // :183-185
    std::vector<ColPtr> rk;
    for (const Expr& e : right_keys) rk.push_back(eval_expr(e, right));
    std::vector<uint64_t> hashes(right.n, 0);
    create_hashes(rk, hashes);
    std::vector<int64_t> li;
    std::vector<uint8_t> li_valid;
    std::vector<uint32_t> ri;
    bool keep_right = join_type == SQLRS_JOIN_RIGHT || join_type == SQLRS_JOIN_FULL;
    for (int64_t row = 0; row < right.n; row++) {  // :225-248
      auto it = left_hashmap.find(hashes[row]);
      bool any = false;
      if (it != left_hashmap.end()) {
        for (int64_t i : it->second) {
          if (opt.match_mode == SQLRS_MATCH_HASH_AND_KEY && !keys_equal(rk, row, i)) continue;
          li.push_back(i);
          li_valid.push_back(1);
          ri.push_back((uint32_t)row);
          any = true;
        }
        if (opt.match_mode == SQLRS_MATCH_HASH_ONLY) any = true;  // `if let Some(indices)` even when the Vec is empty
      }
      if (!any && keep_right) {
        li.push_back(0);
        li_valid.push_back(0);
        ri.push_back((uint32_t)row);
      }
    }
    if (!filter.empty()) {  // apply_join_filter, :47-127
      Batch inter = build_batch(right, li, li_valid, ri);
      ColPtr mask = eval_expr(filter, inter);
      if (mask->dtype != SQLRS_DT_BOOL) fail(SQLRS_ERR_INTERNAL, "join filter expected evaluate boolean array");
      std::vector<int64_t> fl;
      std::vector<uint8_t> flv;
      std::vector<uint32_t> fr;
      for (int64_t k = 0; k < inter.n; k++)
        if (mask->is_valid(k) && mask->i[k]) {
          fl.push_back(li[k]);
          flv.push_back(li_valid[k]);
          fr.push_back(ri[k]);
        }
      if (keep_right) {  // :73-121
        std::vector<uint8_t> visited_right(right.n, 0);
        for (uint32_t x : fr) visited_right[x] = 1;
        for (int64_t v = 0; v < right.n; v++)
          if (!visited_right[v]) {
            fl.push_back(0);
            flv.push_back(0);
            fr.push_back((uint32_t)v);
          }
      }
      li.swap(fl);
      li_valid.swap(flv);
      ri.swap(fr);
    }
    if (!visited_left.empty() || join_type == SQLRS_JOIN_LEFT || join_type == SQLRS_JOIN_FULL)
      for (size_t k = 0; k < li.size(); k++)
        if (li_valid[k]) visited_left[li[k]] = 1;  // :274-282
    *result = build_batch(right, li, li_valid, ri);  // :284-291
    return true;
  }

  // Left/Full tail, :296-322
  bool finish(Batch* result) {
    seal();
    if (left_batches.empty()) return false;
    if (!(join_type == SQLRS_JOIN_LEFT || join_type == SQLRS_JOIN_FULL)) return false;
    std::vector<int64_t> idx;
    for (int64_t v = 0; v < left_single.n; v++)
      if (!visited_left[v]) idx.push_back(v);
    Batch out;
    out.fields = out_fields;
    out.n = (int64_t)idx.size();
    for (const ColPtr& c : left_single.cols) out.cols.push_back(take(*c, idx));
    for (size_t c = left_single.cols.size(); c < out_fields.size(); c++)
      out.cols.push_back(null_column(out_fields[c].dtype, out.n));
    check_schema(out);
    *result = out;
    return true;
  }
};

}  // namespace oracle
