// ORACLE — TEST INFRASTRUCTURE ONLY.
// The C ABI of include/sqlrs_b200.h compiled with SQLRS_ORACLE_BUILD (symbols sqlrs_oracle_*),
// backed by the CPU restatement in ops.hpp.  Built into oracle/liboracle.so by oracle/Makefile.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
//
// Parity status: the reference cannot be compiled here (no cargo/rustc; it
// needs nightly-2022-07-29 + arrow `simd`), so this restatement is pinned by the reference's own
// golden tests instead — hash KAT (hash_utils.rs:229-247), HashAgg two-chunk test
// (hash_agg.rs:182-222), the 8 HashJoin tables (hash_join.rs:423-750), the executor e2e tests
// (executor/mod.rs:271-396) and the v1 .slt files over tests/csv — see tests/test_reference_golden.py.
// The operators that follow the hot path (Project / Order / Limit, tail_ops.hpp) are pinned by limit.rs:96-101,
// executor/mod.rs:353-396, tests/slt/{order,limit}.slt; DISTINCT aggregates by tests/slt/distinct.slt.
// UNPINNED by any reference test: COUNT over >1 batch (K1), float SUM order, 64-bit collisions (K2),
// NULL-key joins (K3), Int32/Boolean/Utf8 hash_one, multi-batch joins, Float64 min/max NaN, the order of rows that
// tie on every sort key, OFFSET without LIMIT over several batches, NULL as an element of a DISTINCT set.
#define SQLRS_ORACLE_BUILD 1
#include <chrono>
#include <deque>
#include <map>
#include <thread>

#include "../include/sqlrs_b200.h"
#include "../include/sqlrs_tpch_spec.h"
#include "arrow_io.hpp"
#include "csv_reader.hpp"
#include "ops.hpp"
#include "tail_ops.hpp"

using namespace oracle;

static thread_local std::string g_last_error;

template <typename F>
static int guarded(F&& f) {
  try {
    f();
    g_last_error.clear();
    return SQLRS_OK;
  } catch (const Error& e) {
    g_last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return SQLRS_ERR_INTERNAL;
  }
}

// take ownership of an input ArrowArray (moved into the callee), import, release
static Batch consume_batch(ArrowArray* array, const ArrowSchema* schema) {
  if (!array || !array->release) fail(SQLRS_ERR_INVALID_ARG, "input ArrowArray is NULL or already released");
  struct Guard {
    ArrowArray* a;
    ~Guard() {
      if (a->release) a->release(a);
    }
  } guard{array};
  return import_batch(array, schema);
}

static std::vector<AggDesc> copy_aggs(const sqlrs_agg_desc* aggs, int32_t n) {
  std::vector<AggDesc> out;
  if (n > 0 && !aggs) fail(SQLRS_ERR_INVALID_ARG, "aggs is NULL");
  for (int32_t k = 0; k < n; k++) out.push_back(copy_agg(aggs[k]));
  return out;
}
static std::vector<Expr> copy_exprs(const sqlrs_expr* e, int32_t n) {
  std::vector<Expr> out;
  if (n > 0 && !e) fail(SQLRS_ERR_INVALID_ARG, "expression array is NULL");
  for (int32_t k = 0; k < n; k++) out.push_back(copy_expr(&e[k]));
  return out;
}
static std::vector<std::string> copy_names(const char* const* names, int32_t n) {
  std::vector<std::string> out;
  for (int32_t k = 0; k < n; k++) out.push_back(names && names[k] ? names[k] : "");
  return out;
}

struct sqlrs_filter {
  Expr predicate;
};
struct sqlrs_simple_agg {
  SimpleAgg impl;
};
struct sqlrs_hash_agg {
  HashAgg impl;
};
struct sqlrs_hash_join {
  HashJoin impl;
};
struct sqlrs_project {
  Project impl;
};
struct sqlrs_cross_join {
  CrossJoin impl;
  std::deque<Batch> queue;
};
struct sqlrs_table {  // InMemoryTable (src/storage/memory.rs:62-123): the batches, as given
  std::vector<Batch> batches;
};
struct sqlrs_order {
  Order impl;
};
struct sqlrs_limit {
  Limit impl;
};
static std::vector<uint8_t> null_names(const char* const* names, int32_t n) {
  std::vector<uint8_t> out;
  for (int32_t k = 0; k < n; k++) out.push_back(!(names && names[k]));
  return out;
}
static std::vector<uint8_t> copy_asc(const int32_t* asc, int32_t n) {
  std::vector<uint8_t> out;
  for (int32_t k = 0; k < n; k++) out.push_back(asc ? (asc[k] != 0) : 1);
  return out;
}

static HashJoin make_join(int32_t join_type, const sqlrs_expr* lk, const sqlrs_expr* rk, int32_t n_keys,
                          const sqlrs_expr* filter, const ArrowSchema* out_schema, const sqlrs_options* options) {
  if (join_type < SQLRS_JOIN_INNER || join_type > SQLRS_JOIN_FULL) fail(SQLRS_ERR_INVALID_ARG, "bad join type");
  if (n_keys < 1) fail(SQLRS_ERR_INTERNAL, "HashJoin must has on condition");
  HashJoin j;
  j.join_type = join_type;
  j.left_keys = copy_exprs(lk, n_keys);
  j.right_keys = copy_exprs(rk, n_keys);
  if (filter && filter->n_nodes > 0) j.filter = copy_expr(filter);
  j.out_fields = import_fields(out_schema);
  j.opt = copy_options(options);
  return j;
}

// ------------------------------------------------------------------ plan
struct PlanNode {
  sqlrs_plan_node raw;
  Expr predicate;
  std::vector<AggDesc> aggs;
  std::vector<Expr> group_by, left_keys, right_keys;
  std::vector<std::string> group_names;
  std::vector<Field> join_fields;
  std::vector<Expr> exprs;
  std::vector<std::string> expr_names;
  std::vector<uint8_t> asc, keep_field;
};
struct sqlrs_plan {
  std::vector<PlanNode> nodes;
  int root = 0;
  Options opt;
  sqlrs_options raw_opt{};
  std::map<int, std::vector<Batch>> tables;
  std::deque<Batch> results;
  std::string description;
  std::unique_ptr<HashAgg> partial;
  int partial_sel = 0;  // table of `partial` the partial-state calls address (DISTINCT: one table per set)

  std::vector<Batch> run(int idx) {
    if (idx < 0 || idx >= (int)nodes.size()) fail(SQLRS_ERR_INVALID_ARG, "plan: child index out of range");
    PlanNode& n = nodes[idx];
    switch (n.raw.kind) {
      case SQLRS_NODE_SCAN: return tables[n.raw.table_slot];
      case SQLRS_NODE_FILTER: {
        std::vector<Batch> out;
        for (const Batch& b : run(n.raw.child0)) out.push_back(filter_batch(n.predicate, b));
        return out;
      }
      case SQLRS_NODE_SIMPLE_AGG: {
        SimpleAgg a(n.aggs, opt);
        for (const Batch& b : run(n.raw.child0)) a.push(b);
        return {a.finish()};
      }
      case SQLRS_NODE_HASH_AGG: {
        HashAgg a(n.aggs, n.group_by, n.group_names, opt);
        for (const Batch& b : run(n.raw.child0)) a.push(b);
        return {a.finish()};
      }
      case SQLRS_NODE_HASH_JOIN: {
        HashJoin j;
        j.join_type = n.raw.join_type;
        j.left_keys = n.left_keys;
        j.right_keys = n.right_keys;
        j.filter = n.predicate;
        j.out_fields = n.join_fields;
        j.opt = opt;
        for (const Batch& b : run(n.raw.child0)) j.build_push(b);
        std::vector<Batch> out;
        for (const Batch& b : run(n.raw.child1)) {
          Batch r;
          if (j.probe(b, &r)) out.push_back(r);
        }
        Batch tail;
        if (j.finish(&tail)) out.push_back(tail);
        return out;
      }
      case SQLRS_NODE_CROSS_JOIN: {
        CrossJoin j;
        j.out_fields = n.join_fields;
        for (const Batch& b : run(n.raw.child0)) j.build_push(b);
        std::vector<Batch> out;
        for (const Batch& b : run(n.raw.child1))
          for (Batch& r : j.probe(b)) out.push_back(std::move(r));
        return out;
      }
      case SQLRS_NODE_PROJECT: {
        Project p{n.exprs, n.expr_names, n.keep_field};
        std::vector<Batch> out;
        for (const Batch& b : run(n.raw.child0)) out.push_back(p.execute(b));
        return out;
      }
      case SQLRS_NODE_ORDER: {
        Order o{n.exprs, n.asc, {}};
        for (const Batch& b : run(n.raw.child0)) o.push(b);
        return {o.finish()};
      }
      case SQLRS_NODE_LIMIT: {
        Limit l;
        l.limit = n.raw.limit;
        l.offset = n.raw.offset;
        std::vector<Batch> out;
        if (l.limit == 0) return out;  // limit.rs:31-33: the child is never polled
        for (const Batch& b : run(n.raw.child0)) {
          Batch r;
          if (l.push(b, &r)) out.push_back(r);
          if (l.done) break;
        }
        return out;
      }
    }
    fail(SQLRS_ERR_INVALID_ARG, "plan: unknown node kind");
  }
};

extern "C" {

int sqlrs_oracle_abi_version(void) { return SQLRS_ABI_VERSION; }
const char* sqlrs_oracle_last_error(void) { return g_last_error.c_str(); }
int64_t sqlrs_oracle_kernel_launches(void) { return 0; }

int sqlrs_oracle_create_hashes(ArrowArray* columns, const ArrowSchema* schema, uint64_t* out_hashes) {
  return guarded([&] {
    Batch b = consume_batch(columns, schema);
    std::vector<uint64_t> h(b.n, 0);
    create_hashes(b.cols, h);
    if (b.n) std::memcpy(out_hashes, h.data(), sizeof(uint64_t) * b.n);
  });
}

int sqlrs_oracle_eval_expr(const sqlrs_expr* expr, const sqlrs_options*, ArrowArray* batch, const ArrowSchema* schema,
                           ArrowArray* out, ArrowSchema* out_schema) {
  return guarded([&] {
    Batch b = consume_batch(batch, schema);
    Expr e = copy_expr(expr);
    ColPtr c = eval_expr(e, b);
    Batch r;
    r.n = b.n;
    r.cols.push_back(c);
    r.fields.push_back(Field{"expr", c->dtype, true});
    export_batch(r, out, out_schema);
  });
}

int sqlrs_oracle_filter_create(const sqlrs_expr* predicate, const sqlrs_options*, sqlrs_filter** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    Expr e = copy_expr(predicate);
    if (e.empty()) fail(SQLRS_ERR_INVALID_ARG, "filter needs a predicate");
    *out = new sqlrs_filter{e};
  });
}
int sqlrs_oracle_filter_execute(sqlrs_filter* f, ArrowArray* batch, const ArrowSchema* schema, ArrowArray* out,
                                ArrowSchema* out_schema) {
  return guarded([&] {
    Batch b = consume_batch(batch, schema);
    export_batch(filter_batch(f->predicate, b), out, out_schema);
  });
}
void sqlrs_oracle_filter_destroy(sqlrs_filter* f) { delete f; }

int sqlrs_oracle_simple_agg_create(const sqlrs_agg_desc* aggs, int32_t n_aggs, const sqlrs_options* options,
                                   sqlrs_simple_agg** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    *out = new sqlrs_simple_agg{SimpleAgg(copy_aggs(aggs, n_aggs), copy_options(options))};
  });
}
int sqlrs_oracle_simple_agg_push(sqlrs_simple_agg* a, ArrowArray* batch, const ArrowSchema* schema) {
  return guarded([&] { a->impl.push(consume_batch(batch, schema)); });
}
int sqlrs_oracle_simple_agg_finish(sqlrs_simple_agg* a, ArrowArray* out, ArrowSchema* out_schema) {
  return guarded([&] { export_batch(a->impl.finish(), out, out_schema); });
}
void sqlrs_oracle_simple_agg_destroy(sqlrs_simple_agg* a) { delete a; }

int sqlrs_oracle_hash_agg_create(const sqlrs_agg_desc* aggs, int32_t n_aggs, const sqlrs_expr* group_by,
                                 const char* const* group_names, int32_t n_group_by, const sqlrs_options* options,
                                 sqlrs_hash_agg** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    *out = new sqlrs_hash_agg{HashAgg(copy_aggs(aggs, n_aggs), copy_exprs(group_by, n_group_by),
                                      copy_names(group_names, n_group_by), copy_options(options))};
  });
}
int sqlrs_oracle_hash_agg_push(sqlrs_hash_agg* a, ArrowArray* batch, const ArrowSchema* schema) {
  return guarded([&] { a->impl.push(consume_batch(batch, schema)); });
}
int sqlrs_oracle_hash_agg_finish(sqlrs_hash_agg* a, ArrowArray* out, ArrowSchema* out_schema) {
  return guarded([&] { export_batch(a->impl.finish(), out, out_schema); });
}
void sqlrs_oracle_hash_agg_destroy(sqlrs_hash_agg* a) { delete a; }

int sqlrs_oracle_hash_join_create(int32_t join_type, const sqlrs_expr* left_keys, const sqlrs_expr* right_keys,
                                  int32_t n_keys, const sqlrs_expr* filter, const ArrowSchema* join_output_schema,
                                  const sqlrs_options* options, sqlrs_hash_join** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    *out = new sqlrs_hash_join{make_join(join_type, left_keys, right_keys, n_keys, filter, join_output_schema, options)};
  });
}
int sqlrs_oracle_hash_join_build_push(sqlrs_hash_join* j, ArrowArray* batch, const ArrowSchema* schema) {
  return guarded([&] { j->impl.build_push(consume_batch(batch, schema)); });
}
int sqlrs_oracle_hash_join_probe(sqlrs_hash_join* j, ArrowArray* batch, const ArrowSchema* schema, ArrowArray* out,
                                 ArrowSchema* out_schema, int32_t* has_batch) {
  return guarded([&] {
    Batch b = consume_batch(batch, schema);
    Batch r;
    bool has = j->impl.probe(b, &r);
    if (has_batch) *has_batch = has;
    if (has) export_batch(r, out, out_schema);
  });
}
int sqlrs_oracle_hash_join_finish(sqlrs_hash_join* j, ArrowArray* out, ArrowSchema* out_schema, int32_t* has_batch) {
  return guarded([&] {
    Batch r;
    bool has = j->impl.finish(&r);
    if (has_batch) *has_batch = has;
    if (has) export_batch(r, out, out_schema);
  });
}
void sqlrs_oracle_hash_join_destroy(sqlrs_hash_join* j) { delete j; }

int sqlrs_oracle_cross_join_create(const ArrowSchema* join_output_schema, const sqlrs_options*, sqlrs_cross_join** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    auto* j = new sqlrs_cross_join();
    std::unique_ptr<sqlrs_cross_join> hold(j);
    j->impl.out_fields = import_fields(join_output_schema);
    *out = hold.release();
  });
}
int sqlrs_oracle_cross_join_build_push(sqlrs_cross_join* j, ArrowArray* batch, const ArrowSchema* schema) {
  return guarded([&] { j->impl.build_push(consume_batch(batch, schema)); });
}
int sqlrs_oracle_cross_join_probe(sqlrs_cross_join* j, ArrowArray* batch, const ArrowSchema* schema) {
  return guarded([&] {
    for (Batch& b : j->impl.probe(consume_batch(batch, schema))) j->queue.push_back(std::move(b));
  });
}
int sqlrs_oracle_cross_join_next(sqlrs_cross_join* j, ArrowArray* out, ArrowSchema* out_schema, int32_t* has_batch) {
  return guarded([&] {
    if (j->queue.empty()) {
      if (has_batch) *has_batch = 0;
      return;
    }
    export_batch(j->queue.front(), out, out_schema);
    j->queue.pop_front();
    if (has_batch) *has_batch = 1;
  });
}
void sqlrs_oracle_cross_join_destroy(sqlrs_cross_join* j) { delete j; }

int sqlrs_oracle_project_create(const sqlrs_expr* exprs, const char* const* names, int32_t n_exprs, const sqlrs_options*,
                                sqlrs_project** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    *out = new sqlrs_project{Project{copy_exprs(exprs, n_exprs), copy_names(names, n_exprs), null_names(names, n_exprs)}};
  });
}
int sqlrs_oracle_project_execute(sqlrs_project* p, ArrowArray* batch, const ArrowSchema* schema, ArrowArray* out,
                                 ArrowSchema* out_schema) {
  return guarded([&] { export_batch(p->impl.execute(consume_batch(batch, schema)), out, out_schema); });
}
void sqlrs_oracle_project_destroy(sqlrs_project* p) { delete p; }

int sqlrs_oracle_order_create(const sqlrs_expr* order_by, const int32_t* asc, int32_t n, const sqlrs_options*, sqlrs_order** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    *out = new sqlrs_order{Order{copy_exprs(order_by, n), copy_asc(asc, n), {}}};
  });
}
int sqlrs_oracle_order_push(sqlrs_order* o, ArrowArray* batch, const ArrowSchema* schema) {
  return guarded([&] { o->impl.push(consume_batch(batch, schema)); });
}
int sqlrs_oracle_order_finish(sqlrs_order* o, ArrowArray* out, ArrowSchema* out_schema) {
  return guarded([&] { export_batch(o->impl.finish(), out, out_schema); });
}
void sqlrs_oracle_order_destroy(sqlrs_order* o) { delete o; }

int sqlrs_oracle_limit_create(int64_t limit, int64_t offset, const sqlrs_options*, sqlrs_limit** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    auto* l = new sqlrs_limit();
    l->impl.limit = limit < 0 ? -1 : limit;
    l->impl.offset = offset < 0 ? -1 : offset;
    *out = l;
  });
}
int sqlrs_oracle_limit_push(sqlrs_limit* l, ArrowArray* batch, const ArrowSchema* schema, ArrowArray* out, ArrowSchema* out_schema,
                            int32_t* has_batch, int32_t* done) {
  return guarded([&] {
    Batch b = consume_batch(batch, schema);
    Batch r;
    const bool has = l->impl.push(b, &r);
    if (has_batch) *has_batch = has;
    if (done) *done = l->impl.done;
    if (has) export_batch(r, out, out_schema);
  });
}
void sqlrs_oracle_limit_destroy(sqlrs_limit* l) { delete l; }

int sqlrs_oracle_plan_create(const sqlrs_plan_node* nodes, int32_t n_nodes, int32_t root, const sqlrs_options* options,
                             sqlrs_plan** out) {
  return guarded([&] {
    if (!out || !nodes || n_nodes < 1 || root < 0 || root >= n_nodes) fail(SQLRS_ERR_INVALID_ARG, "bad plan");
    auto* p = new sqlrs_plan();
    std::unique_ptr<sqlrs_plan> hold(p);
    p->opt = copy_options(options);
    p->root = root;
    for (int32_t k = 0; k < n_nodes; k++) {
      PlanNode n;
      n.raw = nodes[k];
      n.predicate = copy_expr(&nodes[k].predicate);
      n.aggs = copy_aggs(nodes[k].aggs, nodes[k].n_aggs);
      n.group_by = copy_exprs(nodes[k].group_by, nodes[k].n_group_by);
      n.group_names = copy_names(nodes[k].group_names, nodes[k].n_group_by);
      if (nodes[k].kind == SQLRS_NODE_HASH_JOIN) {
        n.left_keys = copy_exprs(nodes[k].left_keys, nodes[k].n_keys);
        n.right_keys = copy_exprs(nodes[k].right_keys, nodes[k].n_keys);
        n.join_fields = import_fields(nodes[k].join_output_schema);
      }
      if (nodes[k].kind == SQLRS_NODE_CROSS_JOIN) n.join_fields = import_fields(nodes[k].join_output_schema);
      if (nodes[k].kind == SQLRS_NODE_PROJECT || nodes[k].kind == SQLRS_NODE_ORDER) {
        n.exprs = copy_exprs(nodes[k].exprs, nodes[k].n_exprs);
        n.expr_names = copy_names(nodes[k].expr_names, nodes[k].n_exprs);
        n.asc = copy_asc(nodes[k].order_asc, nodes[k].n_exprs);
        n.keep_field = null_names(nodes[k].expr_names, nodes[k].n_exprs);
      }
      p->nodes.push_back(std::move(n));
    }
    p->description = "oracle: operator-at-a-time CPU restatement";
    *out = hold.release();
  });
}
int sqlrs_oracle_plan_push_table(sqlrs_plan* p, int32_t table_slot, ArrowArray* batch, const ArrowSchema* schema) {
  return guarded([&] { p->tables[table_slot].push_back(consume_batch(batch, schema)); });
}
// one host batch handed to the executor in slices of batch_rows rows — what the reference's scan does with its 1024-row
// batches (src/storage/csv.rs:105, memory.rs:151-170); the slicing happens here, outside any timed executor run
int sqlrs_oracle_plan_push_table_batched(sqlrs_plan* p, int32_t table_slot, ArrowArray* batch, const ArrowSchema* schema, int64_t batch_rows) {
  return guarded([&] {
    if (batch_rows <= 0) fail(SQLRS_ERR_INVALID_ARG, "push_table_batched: batch_rows must be positive");
    Batch whole = consume_batch(batch, schema);
    if (whole.n == 0) {
      p->tables[table_slot].push_back(std::move(whole));
      return;
    }
    for (int64_t off = 0; off < whole.n; off += batch_rows) {
      Batch b;
      b.fields = whole.fields;
      b.n = std::min(batch_rows, whole.n - off);
      for (const ColPtr& c : whole.cols) b.cols.push_back(slice_column(*c, off, b.n));
      p->tables[table_slot].push_back(std::move(b));
    }
  });
}
int sqlrs_oracle_plan_result_shape(sqlrs_plan*, int64_t*, int32_t*, int32_t*) {
  g_last_error = "the oracle keeps no device results";
  return SQLRS_ERR_UNSUPPORTED;
}
int sqlrs_oracle_plan_next_to_device(sqlrs_plan*, void* const*, int32_t) {
  g_last_error = "the oracle keeps no device results";
  return SQLRS_ERR_UNSUPPORTED;
}
int sqlrs_oracle_plan_export_partials_partitioned(sqlrs_plan*, void*, int32_t, int64_t, int64_t*) {
  g_last_error = "the oracle exchanges partials as host batches only";
  return SQLRS_ERR_UNSUPPORTED;
}
int sqlrs_oracle_kernel_events_collect(char** json_out) {
  if (json_out) {
    *json_out = (char*)std::malloc(3);
    if (*json_out) std::memcpy(*json_out, "{}", 3);
  }
  return SQLRS_OK;
}
int sqlrs_oracle_plan_push_table_device(sqlrs_plan*, int32_t, ArrowDeviceArray*, const ArrowSchema*) {
  g_last_error = "the oracle takes host batches only";
  return SQLRS_ERR_UNSUPPORTED;
}
int sqlrs_oracle_table_create(const sqlrs_options*, sqlrs_table** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    *out = new sqlrs_table();
  });
}
int sqlrs_oracle_table_append(sqlrs_table* t, ArrowArray* batch, const ArrowSchema* schema) {
  return guarded([&] {
    if (!t) fail(SQLRS_ERR_INVALID_ARG, "table is NULL");
    Batch b = consume_batch(batch, schema);
    if (!t->batches.empty() && t->batches[0].fields.size() != b.fields.size()) fail(SQLRS_ERR_ARROW, "table_append: schema mismatch");
    t->batches.push_back(std::move(b));
  });
}
int sqlrs_oracle_table_read_csv(const char* path, int32_t has_header, int32_t delimiter, int64_t batch_rows, int64_t bounds_offset, int64_t bounds_limit,
                                const int32_t* projection, int32_t n_projection, const sqlrs_options*, sqlrs_table** out) {
  return guarded([&] {
    if (!out || !path) fail(SQLRS_ERR_INVALID_ARG, "path / out is NULL");
    std::vector<int> proj;
    for (int32_t k = 0; projection && k < n_projection; k++) proj.push_back(projection[k]);
    auto t = std::make_unique<sqlrs_table>();
    t->batches = read_csv(path, has_header != 0, (char)delimiter, batch_rows > 0 ? batch_rows : 1024, bounds_offset, bounds_limit, proj);
    *out = t.release();
  });
}
int64_t sqlrs_oracle_table_num_rows(sqlrs_table* t) {
  int64_t n = 0;
  if (t)
    for (const Batch& b : t->batches) n += b.n;
  return n;
}
int32_t sqlrs_oracle_table_num_batches(sqlrs_table* t) { return t ? (int32_t)t->batches.size() : 0; }
int sqlrs_oracle_table_read(sqlrs_table* t, int32_t batch_index, const int32_t* projection, int32_t n_projection, ArrowArray* out,
                            ArrowSchema* out_schema, int32_t* has_batch) {
  return guarded([&] {
    if (!t) fail(SQLRS_ERR_INVALID_ARG, "table is NULL");
    if (batch_index < 0 || batch_index >= (int32_t)t->batches.size()) {  // next_batch past the end: None (memory.rs:159-168)
      if (has_batch) *has_batch = 0;
      return;
    }
    const Batch& b = t->batches[(size_t)batch_index];
    Batch r;
    r.n = b.n;
    if (projection) {
      for (int32_t k = 0; k < n_projection; k++) {
        if (projection[k] < 0 || projection[k] >= (int32_t)b.cols.size()) fail(SQLRS_ERR_INVALID_ARG, "table_read: projection index out of range");
        r.fields.push_back(b.fields[(size_t)projection[k]]);
        r.cols.push_back(b.cols[(size_t)projection[k]]);
      }
    } else {
      r = b;
    }
    export_batch(r, out, out_schema);
    if (has_batch) *has_batch = 1;
  });
}
void sqlrs_oracle_table_destroy(sqlrs_table* t) { delete t; }
int sqlrs_oracle_plan_push_table_resident(sqlrs_plan* p, int32_t table_slot, sqlrs_table* t) {
  return guarded([&] {
    if (!p || !t) fail(SQLRS_ERR_INVALID_ARG, "plan / table is NULL");
    for (const Batch& b : t->batches) p->tables[table_slot].push_back(b);
  });
}
int sqlrs_oracle_plan_execute(sqlrs_plan* p) {
  return guarded([&] {
    p->results.clear();
    for (Batch& b : p->run(p->root)) p->results.push_back(std::move(b));
  });
}
int sqlrs_oracle_plan_next(sqlrs_plan* p, ArrowArray* out, ArrowSchema* out_schema, int32_t* has_batch) {
  return guarded([&] {
    if (p->results.empty()) {
      if (has_batch) *has_batch = 0;
      return;
    }
    export_batch(p->results.front(), out, out_schema);
    p->results.pop_front();
    if (has_batch) *has_batch = 1;
  });
}
int sqlrs_oracle_plan_reset(sqlrs_plan* p) {
  return guarded([&] {
    p->tables.clear();
    p->results.clear();
  });
}
int sqlrs_oracle_plan_clear_table(sqlrs_plan* p, int32_t table_slot) {
  return guarded([&] { p->tables.erase(table_slot); });
}
const char* sqlrs_oracle_plan_describe(sqlrs_plan* p) { return p->description.c_str(); }
void sqlrs_oracle_plan_destroy(sqlrs_plan* p) { delete p; }
int sqlrs_oracle_plan_execute_partial(sqlrs_plan* p, int64_t row_base) {
  return guarded([&] {
    PlanNode& n = p->nodes[p->root];
    if (n.raw.kind != SQLRS_NODE_HASH_AGG) fail(SQLRS_ERR_UNSUPPORTED, "oracle: execute_partial needs a HashAgg root");
    p->partial.reset(new HashAgg(n.aggs, n.group_by, n.group_names, p->opt));
    p->partial->rows_seen = row_base;
    p->partial_sel = 0;
    for (const Batch& b : p->run(n.raw.child0)) p->partial->push(b);
  });
}
int sqlrs_oracle_plan_partials_tables(sqlrs_plan* p, int32_t* n_tables) {
  return guarded([&] {
    if (!p->partial) fail(SQLRS_ERR_INVALID_ARG, "partials_tables before execute_partial");
    *n_tables = p->partial->partial_tables();
  });
}
int sqlrs_oracle_plan_select_partials_table(sqlrs_plan* p, int32_t index) {
  return guarded([&] {
    if (!p->partial) fail(SQLRS_ERR_INVALID_ARG, "select_partials_table before execute_partial");
    if (index < 0 || index >= p->partial->partial_tables()) fail(SQLRS_ERR_INVALID_ARG, "partials table index out of range");
    p->partial_sel = index;
  });
}
int sqlrs_oracle_plan_export_partials(sqlrs_plan* p, ArrowArray* out, ArrowSchema* out_schema) {
  return guarded([&] {
    if (!p->partial) fail(SQLRS_ERR_INVALID_ARG, "export_partials before execute_partial");
    export_batch(p->partial->export_partials(p->partial_sel), out, out_schema);
  });
}
int sqlrs_oracle_plan_clear_partials(sqlrs_plan* p) {
  return guarded([&] {
    if (!p->partial) fail(SQLRS_ERR_INVALID_ARG, "clear_partials before execute_partial");
    p->partial->clear_partials(p->partial_sel);
  });
}
int sqlrs_oracle_plan_merge_partials(sqlrs_plan* p, ArrowArray* partials, const ArrowSchema* schema) {
  return guarded([&] {
    if (!p->partial) fail(SQLRS_ERR_INVALID_ARG, "merge_partials before execute_partial");
    p->partial->merge_partials(consume_batch(partials, schema), p->partial_sel);
  });
}
int sqlrs_oracle_plan_finish_partial(sqlrs_plan* p) {
  return guarded([&] {
    if (!p->partial) fail(SQLRS_ERR_INVALID_ARG, "finish before execute_partial");
    p->results.push_back(p->partial->finish());
    p->partial.reset();
  });
}
int sqlrs_oracle_plan_partials_row_words(sqlrs_plan*, int32_t*) {
  g_last_error = "the oracle exchanges partials as host batches only";
  return SQLRS_ERR_UNSUPPORTED;
}
int sqlrs_oracle_plan_export_partials_device(sqlrs_plan*, void*, int64_t) {
  g_last_error = "the oracle exchanges partials as host batches only";
  return SQLRS_ERR_UNSUPPORTED;
}
int sqlrs_oracle_plan_merge_partials_device(sqlrs_plan*, const void*, int32_t, int64_t) {
  g_last_error = "the oracle exchanges partials as host batches only";
  return SQLRS_ERR_UNSUPPORTED;
}
double sqlrs_oracle_plan_scan_kernel_ms(sqlrs_plan*, int64_t* n_launches) {
  if (n_launches) *n_launches = 0;
  return 0.0;
}

// ------------------------------------------------------------------ synthetic tables
int32_t sqlrs_oracle_tpch_num_columns(int32_t table) {
  switch (table) {
    case SQLRS_TPCH_CUSTOMER: return SQLRS_CUSTOMER_NCOLS;
    case SQLRS_TPCH_ORDERS: return SQLRS_ORDERS_NCOLS;
    case SQLRS_TPCH_LINEITEM: return SQLRS_LINEITEM_NCOLS;
  }
  return -1;
}
int64_t sqlrs_oracle_tpch_num_rows(const sqlrs_tpch_dims* dims, int32_t table) {
  if (!dims) return -1;
  switch (table) {
    case SQLRS_TPCH_CUSTOMER: return dims->n_customer;
    case SQLRS_TPCH_ORDERS: return dims->n_orders;
    case SQLRS_TPCH_LINEITEM: return sqlrs_tpch_lineitem_rows(dims->n_orders);
  }
  return -1;
}
int sqlrs_oracle_tpch_generate(const sqlrs_tpch_dims* dims, int32_t table, int64_t row_begin, int64_t row_end,
                               void* const* columns, void*) {
  return guarded([&] {
    int32_t ncols = sqlrs_oracle_tpch_num_columns(table);
    int64_t nrows = sqlrs_oracle_tpch_num_rows(dims, table);
    if (ncols < 0 || nrows < 0) fail(SQLRS_ERR_INVALID_ARG, "bad table / dims");
    if (row_begin < 0 || row_end < row_begin || row_end > nrows) fail(SQLRS_ERR_INVALID_ARG, "row range out of bounds");
    for (int32_t c = 0; c < ncols; c++) {
      uint64_t* dst = (uint64_t*)columns[c];
      if (!dst) continue;  // NULL = column not wanted
      // input production only (never timed): split the row range over the host threads
      int64_t total = row_end - row_begin;
      int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::thread::hardware_concurrency(), total / 65536));
      std::vector<std::thread> pool;
      for (int t = 0; t < nt; t++) {
        int64_t lo = row_begin + total * t / nt, hi = row_begin + total * (t + 1) / nt;
        pool.emplace_back([=] {
          for (int64_t r = lo; r < hi; r++)
            dst[r - row_begin] = sqlrs_tpch_cell(table, c, r, dims->n_customer, dims->flags_mode);
        });
      }
      for (auto& th : pool) th.join();
    }
  });
}

int sqlrs_oracle_debug_compile_agg(const sqlrs_agg_desc*, int32_t, const sqlrs_expr*, int32_t, const sqlrs_expr*, const ArrowSchema*,
                                   const sqlrs_options*, int32_t, char**) {
  g_last_error = "the oracle generates no kernels";
  return SQLRS_ERR_UNSUPPORTED;
}
int sqlrs_oracle_debug_compile_joinagg(const sqlrs_agg_desc*, int32_t, const sqlrs_expr*, int32_t, const sqlrs_expr*, int32_t, const sqlrs_expr*,
                                       const sqlrs_expr*, const ArrowSchema*, const ArrowSchema*, const sqlrs_options*, int32_t, char**) {
  g_last_error = "the oracle generates no kernels";
  return SQLRS_ERR_UNSUPPORTED;
}
int sqlrs_oracle_debug_compile_joinprobe(const sqlrs_expr*, int32_t, const sqlrs_expr*, const ArrowSchema*, const sqlrs_options*, int32_t, char**) {
  g_last_error = "the oracle generates no kernels";
  return SQLRS_ERR_UNSUPPORTED;
}
int sqlrs_oracle_debug_compile_joinchain(const sqlrs_expr*, int32_t, const sqlrs_expr*, const sqlrs_expr*, const ArrowSchema*, const ArrowSchema*,
                                         const sqlrs_options*, int32_t, char**) {
  g_last_error = "the oracle generates no kernels";
  return SQLRS_ERR_UNSUPPORTED;
}
int sqlrs_oracle_debug_compile_eval(const sqlrs_expr*, int32_t, int32_t, const ArrowSchema*, int32_t, char**) {
  g_last_error = "the oracle generates no kernels";
  return SQLRS_ERR_UNSUPPORTED;
}
void sqlrs_oracle_free(void* p) { std::free(p); }

}  // extern "C"
