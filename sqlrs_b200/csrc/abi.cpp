// sqlrs_b200 — the extern "C" boundary of include/sqlrs_b200.h (symbols sqlrs_*), CUDA build.
// Every entry point catches C++ exceptions and maps them to the status codes of the header; the
// library never aborts the process (the reference panics in several of these places).
#include <deque>

#include "../../include/sqlrs_tpch_spec.h"
#include "join.hpp"
#include "kernels_aot.hpp"
#include "ops.hpp"
#include "csv.hpp"
#include "plan.hpp"
#include "tail.hpp"

using namespace sq;

static thread_local std::string g_last_error;

template <typename F>
static int guarded(F&& f) {
  try {
    f();
    g_last_error.clear();
    return SQLRS_OK;
  } catch (const Error& e) {
    g_last_error = e.what();
    cudaGetLastError();
    return e.code;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    cudaGetLastError();
    return SQLRS_ERR_INTERNAL;
  }
}

static std::vector<ExprCopy> copy_exprs(const sqlrs_expr* e, int32_t n) {
  std::vector<ExprCopy> out;
  if (n > 0 && !e) fail(SQLRS_ERR_INVALID_ARG, "expression array is NULL");
  for (int32_t k = 0; k < n; k++) out.push_back(copy_expr(&e[k]));
  return out;
}
static std::vector<bool> null_names(const char* const* names, int32_t n) {
  std::vector<bool> out;
  for (int32_t k = 0; k < n; k++) out.push_back(!(names && names[k]));
  return out;
}
static std::vector<std::string> copy_names(const char* const* names, int32_t n) {
  std::vector<std::string> out;
  for (int32_t k = 0; k < n; k++) out.push_back(names && names[k] ? names[k] : "");
  return out;
}

struct sqlrs_filter {
  FilterOp op;
  sqlrs_filter(const ExprCopy& p, const Options& o) : op(p, o) {}
};
struct sqlrs_simple_agg {
  AggOp op;
  sqlrs_simple_agg(std::vector<AggSpec> a, const Options& o) : op(std::move(a), {}, {}, true, {}, o) {}
};
struct sqlrs_hash_agg {
  AggOp op;
  sqlrs_hash_agg(std::vector<AggSpec> a, std::vector<ExprCopy> g, std::vector<std::string> names, const Options& o)
      : op(std::move(a), std::move(g), std::move(names), false, {}, o) {}
};
struct sqlrs_hash_join {
  JoinOp op;
  sqlrs_hash_join(int type, std::vector<ExprCopy> lk, std::vector<ExprCopy> rk, ExprCopy filter, std::vector<Field> fields,
                  const Options& o)
      : op(type, std::move(lk), std::move(rk), std::move(filter), std::move(fields), o) {}
};
struct sqlrs_project {
  ProjectOp op;
  sqlrs_project(std::vector<ExprCopy> e, std::vector<std::string> names, std::vector<bool> keep, const Options& o)
      : op(std::move(e), std::move(names), std::move(keep), o) {}
};
struct sqlrs_order {
  OrderOp op;
  sqlrs_order(std::vector<ExprCopy> e, std::vector<bool> asc, const Options& o) : op(std::move(e), std::move(asc), o) {}
};
struct sqlrs_limit {
  LimitOp op;
  sqlrs_limit(int64_t limit, int64_t offset, const Options& o) : op(limit, offset, o) {}
};
struct sqlrs_cross_join {
  CrossJoinOp op;
  std::deque<DBatch> queue;
  sqlrs_cross_join(std::vector<Field> f, const Options& o) : op(std::move(f), o) {}
};
struct sqlrs_table {  // batches resident in HBM (columns shared with every plan they were pushed to)
  Ctx ctx;
  std::vector<DBatch> batches;
  explicit sqlrs_table(const Options& o) : ctx(o) {}
};
struct sqlrs_plan {
  Plan impl;
  sqlrs_plan(const sqlrs_plan_node* nodes, int32_t n, int32_t root, const Options& o) : impl(nodes, n, root, o) {}
};

extern "C" {

int sqlrs_abi_version(void) { return SQLRS_ABI_VERSION; }
const char* sqlrs_last_error(void) { return g_last_error.c_str(); }
int64_t sqlrs_kernel_launches(void) { return g_kernel_launches.load(); }

int sqlrs_create_hashes(ArrowArray* columns, const ArrowSchema* schema, uint64_t* out_hashes) {
  return guarded([&] {
    Options opt;
    Ctx ctx(opt);
    DBatch b = import_batch_host(ctx, columns, schema);
    EvalRequest req;
    for (size_t c = 0; c < b.cols.size(); c++) {
      ExprNodeCopy n;
      n.op = SQLRS_OP_INPUT_REF;
      n.index = (int)c;
      n.dtype = b.cols[c].dtype;
      req.exprs.push_back(ExprCopy{n});
      req.is_key.push_back(true);
    }
    req.outs.push_back({OUT_HASH, 0});
    EvalProgram prog(std::move(req));
    EvalResult r = prog.run(ctx, b, "create_hashes");
    if (b.n) SQ_CUDA(cudaMemcpyAsync(out_hashes, r.cols[0].data, (size_t)b.n * 8, cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
  });
}

int sqlrs_eval_expr(const sqlrs_expr* expr, const sqlrs_options* options, ArrowArray* batch, const ArrowSchema* schema, ArrowArray* out,
                    ArrowSchema* out_schema) {
  return guarded([&] {
    Ctx ctx(copy_options(options));
    DBatch b = import_batch_host(ctx, batch, schema);
    EvalRequest req;
    req.exprs.push_back(copy_expr(expr));
    req.is_key.push_back(false);
    req.outs.push_back({OUT_VALUE, 0});
    EvalProgram prog(std::move(req));
    EvalResult r = prog.run(ctx, b, "expression");
    DBatch res;
    res.n = b.n;
    res.cols.push_back(r.cols[0]);
    res.fields.push_back(Field{"expr", r.cols[0].dtype, true});
    export_batch_host(ctx, res, out, out_schema);
  });
}

int sqlrs_filter_create(const sqlrs_expr* predicate, const sqlrs_options* options, sqlrs_filter** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    *out = new sqlrs_filter(copy_expr(predicate), copy_options(options));
  });
}
int sqlrs_filter_execute(sqlrs_filter* f, ArrowArray* batch, const ArrowSchema* schema, ArrowArray* out, ArrowSchema* out_schema) {
  return guarded([&] {
    if (!f) fail(SQLRS_ERR_INVALID_ARG, "filter handle is NULL");
    f->op.ctx().activate();
    DBatch b = import_batch_host(f->op.ctx(), batch, schema);
    DBatch r = f->op.execute(b);
    export_batch_host(f->op.ctx(), r, out, out_schema);
  });
}
void sqlrs_filter_destroy(sqlrs_filter* f) { delete f; }

int sqlrs_simple_agg_create(const sqlrs_agg_desc* aggs, int32_t n_aggs, const sqlrs_options* options, sqlrs_simple_agg** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    *out = new sqlrs_simple_agg(copy_aggs(aggs, n_aggs), copy_options(options));
  });
}
int sqlrs_simple_agg_push(sqlrs_simple_agg* a, ArrowArray* batch, const ArrowSchema* schema) {
  return guarded([&] {
    if (!a) fail(SQLRS_ERR_INVALID_ARG, "handle is NULL");
    a->op.ctx().activate();
    a->op.push(import_batch_host(a->op.ctx(), batch, schema));
  });
}
int sqlrs_simple_agg_finish(sqlrs_simple_agg* a, ArrowArray* out, ArrowSchema* out_schema) {
  return guarded([&] {
    if (!a) fail(SQLRS_ERR_INVALID_ARG, "handle is NULL");
    a->op.finish_host(out, out_schema);
  });
}
void sqlrs_simple_agg_destroy(sqlrs_simple_agg* a) { delete a; }

int sqlrs_hash_agg_create(const sqlrs_agg_desc* aggs, int32_t n_aggs, const sqlrs_expr* group_by, const char* const* group_names,
                          int32_t n_group_by, const sqlrs_options* options, sqlrs_hash_agg** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    *out = new sqlrs_hash_agg(copy_aggs(aggs, n_aggs), copy_exprs(group_by, n_group_by), copy_names(group_names, n_group_by),
                              copy_options(options));
  });
}
int sqlrs_hash_agg_push(sqlrs_hash_agg* a, ArrowArray* batch, const ArrowSchema* schema) {
  return guarded([&] {
    if (!a) fail(SQLRS_ERR_INVALID_ARG, "handle is NULL");
    a->op.ctx().activate();
    a->op.push(import_batch_host(a->op.ctx(), batch, schema));
  });
}
int sqlrs_hash_agg_finish(sqlrs_hash_agg* a, ArrowArray* out, ArrowSchema* out_schema) {
  return guarded([&] {
    if (!a) fail(SQLRS_ERR_INVALID_ARG, "handle is NULL");
    a->op.finish_host(out, out_schema);
  });
}
void sqlrs_hash_agg_destroy(sqlrs_hash_agg* a) { delete a; }

int sqlrs_hash_join_create(int32_t join_type, const sqlrs_expr* left_keys, const sqlrs_expr* right_keys, int32_t n_keys,
                           const sqlrs_expr* filter, const ArrowSchema* join_output_schema, const sqlrs_options* options,
                           sqlrs_hash_join** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    if (join_type < SQLRS_JOIN_INNER || join_type > SQLRS_JOIN_FULL) fail(SQLRS_ERR_INVALID_ARG, "bad join type");
    if (n_keys < 1) fail(SQLRS_ERR_INTERNAL, "HashJoin must has on condition");
    ExprCopy flt;
    if (filter && filter->n_nodes > 0) flt = copy_expr(filter);
    *out = new sqlrs_hash_join(join_type, copy_exprs(left_keys, n_keys), copy_exprs(right_keys, n_keys), flt,
                               import_fields(join_output_schema), copy_options(options));
  });
}
int sqlrs_hash_join_build_push(sqlrs_hash_join* j, ArrowArray* batch, const ArrowSchema* schema) {
  return guarded([&] {
    if (!j) fail(SQLRS_ERR_INVALID_ARG, "handle is NULL");
    j->op.ctx().activate();
    j->op.build_push(import_batch_host(j->op.ctx(), batch, schema));
  });
}
int sqlrs_hash_join_probe(sqlrs_hash_join* j, ArrowArray* batch, const ArrowSchema* schema, ArrowArray* out, ArrowSchema* out_schema,
                          int32_t* has_batch) {
  return guarded([&] {
    if (!j) fail(SQLRS_ERR_INVALID_ARG, "handle is NULL");
    j->op.ctx().activate();
    DBatch b = import_batch_host(j->op.ctx(), batch, schema);
    DBatch r;
    bool has = j->op.probe(b, &r);
    if (has_batch) *has_batch = has;
    if (has) export_batch_host(j->op.ctx(), r, out, out_schema);
    else j->op.ctx().sync();
  });
}
int sqlrs_hash_join_finish(sqlrs_hash_join* j, ArrowArray* out, ArrowSchema* out_schema, int32_t* has_batch) {
  return guarded([&] {
    if (!j) fail(SQLRS_ERR_INVALID_ARG, "handle is NULL");
    j->op.ctx().activate();
    DBatch r;
    bool has = j->op.finish(&r);
    if (has_batch) *has_batch = has;
    if (has) export_batch_host(j->op.ctx(), r, out, out_schema);
  });
}
void sqlrs_hash_join_destroy(sqlrs_hash_join* j) { delete j; }

int sqlrs_cross_join_create(const ArrowSchema* join_output_schema, const sqlrs_options* options, sqlrs_cross_join** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    *out = new sqlrs_cross_join(import_fields(join_output_schema), copy_options(options));
  });
}
int sqlrs_cross_join_build_push(sqlrs_cross_join* j, ArrowArray* batch, const ArrowSchema* schema) {
  return guarded([&] {
    if (!j) fail(SQLRS_ERR_INVALID_ARG, "handle is NULL");
    j->op.ctx().activate();
    j->op.build_push(import_batch_host(j->op.ctx(), batch, schema));
  });
}
int sqlrs_cross_join_probe(sqlrs_cross_join* j, ArrowArray* batch, const ArrowSchema* schema) {
  return guarded([&] {
    if (!j) fail(SQLRS_ERR_INVALID_ARG, "handle is NULL");
    j->op.ctx().activate();
    DBatch b = import_batch_host(j->op.ctx(), batch, schema);
    for (DBatch& r : j->op.probe(b)) j->queue.push_back(std::move(r));
  });
}
int sqlrs_cross_join_next(sqlrs_cross_join* j, ArrowArray* out, ArrowSchema* out_schema, int32_t* has_batch) {
  return guarded([&] {
    if (!j) fail(SQLRS_ERR_INVALID_ARG, "handle is NULL");
    if (j->queue.empty()) {
      if (has_batch) *has_batch = 0;
      return;
    }
    j->op.ctx().activate();
    export_batch_host(j->op.ctx(), j->queue.front(), out, out_schema);
    j->queue.pop_front();
    if (has_batch) *has_batch = 1;
  });
}
void sqlrs_cross_join_destroy(sqlrs_cross_join* j) { delete j; }

int sqlrs_project_create(const sqlrs_expr* exprs, const char* const* names, int32_t n_exprs, const sqlrs_options* options,
                         sqlrs_project** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    *out = new sqlrs_project(copy_exprs(exprs, n_exprs), copy_names(names, n_exprs), null_names(names, n_exprs), copy_options(options));
  });
}
int sqlrs_project_execute(sqlrs_project* p, ArrowArray* batch, const ArrowSchema* schema, ArrowArray* out, ArrowSchema* out_schema) {
  return guarded([&] {
    if (!p) fail(SQLRS_ERR_INVALID_ARG, "handle is NULL");
    p->op.ctx().activate();
    DBatch b = import_batch_host(p->op.ctx(), batch, schema);
    DBatch r = p->op.execute(b);
    export_batch_host(p->op.ctx(), r, out, out_schema);
  });
}
void sqlrs_project_destroy(sqlrs_project* p) { delete p; }

int sqlrs_order_create(const sqlrs_expr* order_by, const int32_t* asc, int32_t n, const sqlrs_options* options, sqlrs_order** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    std::vector<bool> dirs;
    for (int32_t k = 0; k < n; k++) dirs.push_back(asc ? asc[k] != 0 : true);
    *out = new sqlrs_order(copy_exprs(order_by, n), dirs, copy_options(options));
  });
}
int sqlrs_order_push(sqlrs_order* o, ArrowArray* batch, const ArrowSchema* schema) {
  return guarded([&] {
    if (!o) fail(SQLRS_ERR_INVALID_ARG, "handle is NULL");
    o->op.ctx().activate();
    o->op.push(import_batch_host(o->op.ctx(), batch, schema));
  });
}
int sqlrs_order_finish(sqlrs_order* o, ArrowArray* out, ArrowSchema* out_schema) {
  return guarded([&] {
    if (!o) fail(SQLRS_ERR_INVALID_ARG, "handle is NULL");
    o->op.ctx().activate();
    DBatch r = o->op.finish();
    export_batch_host(o->op.ctx(), r, out, out_schema);
  });
}
void sqlrs_order_destroy(sqlrs_order* o) { delete o; }

int sqlrs_limit_create(int64_t limit, int64_t offset, const sqlrs_options* options, sqlrs_limit** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    *out = new sqlrs_limit(limit, offset, copy_options(options));
  });
}
int sqlrs_limit_push(sqlrs_limit* l, ArrowArray* batch, const ArrowSchema* schema, ArrowArray* out, ArrowSchema* out_schema,
                     int32_t* has_batch, int32_t* done) {
  return guarded([&] {
    if (!l) fail(SQLRS_ERR_INVALID_ARG, "handle is NULL");
    l->op.ctx().activate();
    DBatch b = import_batch_host(l->op.ctx(), batch, schema);
    DBatch r;
    const bool has = l->op.push(b, &r);
    if (has_batch) *has_batch = has;
    if (done) *done = l->op.done();
    if (has) export_batch_host(l->op.ctx(), r, out, out_schema);
    else l->op.ctx().sync();
  });
}
void sqlrs_limit_destroy(sqlrs_limit* l) { delete l; }

int sqlrs_plan_create(const sqlrs_plan_node* nodes, int32_t n_nodes, int32_t root, const sqlrs_options* options, sqlrs_plan** out) {
  return guarded([&] {
    if (!out || !nodes || n_nodes < 1 || root < 0 || root >= n_nodes) fail(SQLRS_ERR_INVALID_ARG, "bad plan");
    *out = new sqlrs_plan(nodes, n_nodes, root, copy_options(options));
  });
}
int sqlrs_plan_push_table(sqlrs_plan* p, int32_t table_slot, ArrowArray* batch, const ArrowSchema* schema) {
  return guarded([&] {
    if (!p) fail(SQLRS_ERR_INVALID_ARG, "plan is NULL");
    p->impl.ctx().activate();
    p->impl.push_table(table_slot, import_batch_host(p->impl.ctx(), batch, schema));
  });
}
static char* dup_string(const std::string& s);
int sqlrs_plan_push_table_batched(sqlrs_plan* p, int32_t table_slot, ArrowArray* batch, const ArrowSchema* schema, int64_t batch_rows) {
  return guarded([&] {
    if (!p) fail(SQLRS_ERR_INVALID_ARG, "plan is NULL");
    p->impl.ctx().activate();
    p->impl.push_table_batched(table_slot, import_batch_host(p->impl.ctx(), batch, schema), batch_rows);
  });
}
int sqlrs_plan_result_shape(sqlrs_plan* p, int64_t* n_rows, int32_t* n_columns, int32_t* has_batch) {
  return guarded([&] {
    if (!p) fail(SQLRS_ERR_INVALID_ARG, "plan is NULL");
    const bool has = p->impl.result_shape(n_rows, n_columns);
    if (has_batch) *has_batch = has ? 1 : 0;
  });
}
int sqlrs_plan_next_to_device(sqlrs_plan* p, void* const* columns, int32_t n_columns) {
  return guarded([&] {
    if (!p || !columns) fail(SQLRS_ERR_INVALID_ARG, "plan / columns is NULL");
    p->impl.ctx().activate();
    p->impl.next_to_device(columns, n_columns);
  });
}
int sqlrs_plan_export_partials_partitioned(sqlrs_plan* p, void* dst, int32_t n_parts, int64_t cap_rows, int64_t* groups_out) {
  return guarded([&] {
    if (!p) fail(SQLRS_ERR_INVALID_ARG, "plan is NULL");
    p->impl.ctx().activate();
    const int64_t g = p->impl.export_partials_partitioned((uint64_t*)dst, n_parts, cap_rows);
    if (groups_out) *groups_out = g;
  });
}
int sqlrs_kernel_events_collect(char** json_out) {
  return guarded([&] {
    if (!json_out) fail(SQLRS_ERR_INVALID_ARG, "json_out is NULL");
    *json_out = dup_string(kernel_events_collect_json());
  });
}
int sqlrs_plan_push_table_device(sqlrs_plan* p, int32_t table_slot, ArrowDeviceArray* batch, const ArrowSchema* schema) {
  return guarded([&] {
    if (!p) fail(SQLRS_ERR_INVALID_ARG, "plan is NULL");
    p->impl.ctx().activate();
    p->impl.push_table(table_slot, import_batch_device(p->impl.ctx(), batch, schema));
  });
}
int sqlrs_table_create(const sqlrs_options* options, sqlrs_table** out) {
  return guarded([&] {
    if (!out) fail(SQLRS_ERR_INVALID_ARG, "out is NULL");
    *out = new sqlrs_table(copy_options(options));
  });
}
int sqlrs_table_append(sqlrs_table* t, ArrowArray* batch, const ArrowSchema* schema) {
  return guarded([&] {
    if (!t) fail(SQLRS_ERR_INVALID_ARG, "table is NULL");
    t->ctx.activate();
    DBatch b = import_batch_host(t->ctx, batch, schema);
    if (!t->batches.empty()) {
      const DBatch& first = t->batches[0];
      bool same = first.cols.size() == b.cols.size();
      for (size_t c = 0; same && c < b.cols.size(); c++) same = first.cols[c].dtype == b.cols[c].dtype;
      if (!same) fail(SQLRS_ERR_ARROW, "table_append: schema mismatch");
    }
    t->ctx.sync();  // the H2D copies are complete: any stream may read the batch from now on
    t->batches.push_back(std::move(b));
  });
}
int sqlrs_table_read_csv(const char* path, int32_t has_header, int32_t delimiter, int64_t batch_rows, int64_t bounds_offset, int64_t bounds_limit,
                         const int32_t* projection, int32_t n_projection, const sqlrs_options* options, sqlrs_table** out) {
  return guarded([&] {
    if (!out || !path) fail(SQLRS_ERR_INVALID_ARG, "path / out is NULL");
    auto t = std::make_unique<sqlrs_table>(copy_options(options));
    t->ctx.activate();
    CsvOptions o;
    o.has_header = has_header != 0;
    o.delimiter = (char)delimiter;
    o.batch_rows = batch_rows > 0 ? batch_rows : 1024;
    o.bounds_offset = bounds_offset;
    o.bounds_limit = bounds_limit;
    for (int32_t k = 0; projection && k < n_projection; k++) o.projection.push_back(projection[k]);
    t->batches = read_csv_device(t->ctx, path, o);
    t->ctx.sync();
    *out = t.release();
  });
}
int64_t sqlrs_table_num_rows(sqlrs_table* t) {
  int64_t n = 0;
  if (t)
    for (const DBatch& b : t->batches) n += b.n;
  return n;
}
int32_t sqlrs_table_num_batches(sqlrs_table* t) { return t ? (int32_t)t->batches.size() : 0; }
int sqlrs_table_read(sqlrs_table* t, int32_t batch_index, const int32_t* projection, int32_t n_projection, ArrowArray* out, ArrowSchema* out_schema,
                     int32_t* has_batch) {
  return guarded([&] {
    if (!t) fail(SQLRS_ERR_INVALID_ARG, "table is NULL");
    if (batch_index < 0 || batch_index >= (int32_t)t->batches.size()) {
      if (has_batch) *has_batch = 0;
      return;
    }
    t->ctx.activate();
    const DBatch& b = t->batches[(size_t)batch_index];
    DBatch r;
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
    export_batch_host(t->ctx, r, out, out_schema);
    if (has_batch) *has_batch = 1;
  });
}
void sqlrs_table_destroy(sqlrs_table* t) { delete t; }
int sqlrs_plan_push_table_resident(sqlrs_plan* p, int32_t table_slot, sqlrs_table* t) {
  return guarded([&] {
    if (!p || !t) fail(SQLRS_ERR_INVALID_ARG, "plan / table is NULL");
    if (p->impl.ctx().device != t->ctx.device) fail(SQLRS_ERR_INVALID_ARG, "plan and table live on different devices");
    for (const DBatch& b : t->batches) p->impl.push_table(table_slot, b);  // zero copy: the columns are shared
  });
}
int sqlrs_plan_execute(sqlrs_plan* p) {
  return guarded([&] {
    if (!p) fail(SQLRS_ERR_INVALID_ARG, "plan is NULL");
    p->impl.ctx().activate();
    p->impl.execute();
  });
}
int sqlrs_plan_next(sqlrs_plan* p, ArrowArray* out, ArrowSchema* out_schema, int32_t* has_batch) {
  return guarded([&] {
    if (!p) fail(SQLRS_ERR_INVALID_ARG, "plan is NULL");
    p->impl.ctx().activate();
    bool has = p->impl.next(out, out_schema);
    if (has_batch) *has_batch = has;
  });
}
int sqlrs_plan_reset(sqlrs_plan* p) {
  return guarded([&] {
    if (!p) fail(SQLRS_ERR_INVALID_ARG, "plan is NULL");
    p->impl.reset();
  });
}
int sqlrs_plan_clear_table(sqlrs_plan* p, int32_t table_slot) {
  return guarded([&] {
    if (!p) fail(SQLRS_ERR_INVALID_ARG, "plan is NULL");
    p->impl.clear_table(table_slot);
  });
}
const char* sqlrs_plan_describe(sqlrs_plan* p) { return p ? p->impl.describe() : ""; }
void sqlrs_plan_destroy(sqlrs_plan* p) { delete p; }
int sqlrs_plan_execute_partial(sqlrs_plan* p, int64_t row_base) {
  return guarded([&] {
    if (!p) fail(SQLRS_ERR_INVALID_ARG, "plan is NULL");
    p->impl.ctx().activate();
    p->impl.execute_partial(row_base);
  });
}
int sqlrs_plan_export_partials(sqlrs_plan* p, ArrowArray* out, ArrowSchema* out_schema) {
  return guarded([&] {
    if (!p) fail(SQLRS_ERR_INVALID_ARG, "plan is NULL");
    p->impl.ctx().activate();
    p->impl.export_partials(out, out_schema);
  });
}
int sqlrs_plan_clear_partials(sqlrs_plan* p) {
  return guarded([&] {
    if (!p) fail(SQLRS_ERR_INVALID_ARG, "plan is NULL");
    p->impl.clear_partials();
  });
}
int sqlrs_plan_merge_partials(sqlrs_plan* p, ArrowArray* partials, const ArrowSchema* schema) {
  return guarded([&] {
    if (!p) fail(SQLRS_ERR_INVALID_ARG, "plan is NULL");
    p->impl.ctx().activate();
    p->impl.merge_partials(import_batch_host(p->impl.ctx(), partials, schema));
  });
}
int sqlrs_plan_partials_tables(sqlrs_plan* p, int32_t* n_tables) {
  return guarded([&] {
    if (!p || !n_tables) fail(SQLRS_ERR_INVALID_ARG, "plan / n_tables is NULL");
    *n_tables = p->impl.partials_tables();
  });
}
int sqlrs_plan_select_partials_table(sqlrs_plan* p, int32_t index) {
  return guarded([&] {
    if (!p) fail(SQLRS_ERR_INVALID_ARG, "plan is NULL");
    p->impl.select_partials_table(index);
  });
}
int sqlrs_plan_partials_row_words(sqlrs_plan* p, int32_t* n_words) {
  return guarded([&] {
    if (!p || !n_words) fail(SQLRS_ERR_INVALID_ARG, "plan / n_words is NULL");
    *n_words = p->impl.partial_row_words();
  });
}
int sqlrs_plan_export_partials_device(sqlrs_plan* p, void* dst, int64_t cap_rows) {
  return guarded([&] {
    if (!p || !dst) fail(SQLRS_ERR_INVALID_ARG, "plan / dst is NULL");
    p->impl.ctx().activate();
    p->impl.export_partials_device((uint64_t*)dst, cap_rows);
  });
}
int sqlrs_plan_merge_partials_device(sqlrs_plan* p, const void* src, int32_t n_buffers, int64_t cap_rows) {
  return guarded([&] {
    if (!p || !src) fail(SQLRS_ERR_INVALID_ARG, "plan / src is NULL");
    p->impl.ctx().activate();
    p->impl.merge_partials_device((const uint64_t*)src, n_buffers, cap_rows);
  });
}
int sqlrs_plan_finish_partial(sqlrs_plan* p) {
  return guarded([&] {
    if (!p) fail(SQLRS_ERR_INVALID_ARG, "plan is NULL");
    p->impl.ctx().activate();
    p->impl.finish_partial();
  });
}
double sqlrs_plan_scan_kernel_ms(sqlrs_plan* p, int64_t* n_launches) {
  if (!p) return 0.0;
  if (n_launches) *n_launches = p->impl.scan_kernel_launches();
  return p->impl.scan_kernel_ms();
}

// ------------------------------------------------------------------ synthetic tables
int32_t sqlrs_tpch_num_columns(int32_t table) {
  switch (table) {
    case SQLRS_TPCH_CUSTOMER: return SQLRS_CUSTOMER_NCOLS;
    case SQLRS_TPCH_ORDERS: return SQLRS_ORDERS_NCOLS;
    case SQLRS_TPCH_LINEITEM: return SQLRS_LINEITEM_NCOLS;
  }
  return -1;
}
int64_t sqlrs_tpch_num_rows(const sqlrs_tpch_dims* dims, int32_t table) {
  if (!dims) return -1;
  switch (table) {
    case SQLRS_TPCH_CUSTOMER: return dims->n_customer;
    case SQLRS_TPCH_ORDERS: return dims->n_orders;
    case SQLRS_TPCH_LINEITEM: return sqlrs_tpch_lineitem_rows(dims->n_orders);
  }
  return -1;
}
int sqlrs_tpch_generate(const sqlrs_tpch_dims* dims, int32_t table, int64_t row_begin, int64_t row_end, void* const* columns,
                        void* stream) {
  return guarded([&] {
    int32_t ncols = sqlrs_tpch_num_columns(table);
    int64_t nrows = sqlrs_tpch_num_rows(dims, table);
    if (ncols < 0 || nrows < 0) fail(SQLRS_ERR_INVALID_ARG, "bad table / dims");
    if (row_begin < 0 || row_end < row_begin || row_end > nrows) fail(SQLRS_ERR_INVALID_ARG, "row range out of bounds");
    for (int32_t c = 0; c < ncols; c++) {
      if (!columns[c]) continue;  // NULL = column not wanted
      launch_tpch_generate(table, c, row_begin, row_end - row_begin, dims->n_customer, dims->flags_mode, (uint64_t*)columns[c],
                           (cudaStream_t)stream);
    }
  });
}

// ------------------------------------------------------------------ diagnostics (no GPU needed)
static std::vector<ColInfo> cols_of_schema(const ArrowSchema* schema) {
  std::vector<ColInfo> cols;
  for (const Field& f : import_fields(schema)) cols.push_back(ColInfo{f.dtype, f.nullable, f.nullable});
  return cols;
}
static char* dup_string(const std::string& s) {
  char* p = (char*)std::malloc(s.size() + 1);
  if (!p) fail(SQLRS_ERR_INTERNAL, "out of host memory");
  std::memcpy(p, s.c_str(), s.size() + 1);
  return p;
}

int sqlrs_debug_compile_agg(const sqlrs_agg_desc* aggs, int32_t n_aggs, const sqlrs_expr* group_by, int32_t n_group_by,
                            const sqlrs_expr* fused_predicate, const ArrowSchema* input_schema, const sqlrs_options* options,
                            int32_t compile, char** source_out) {
  return guarded([&] {
    Options opt = copy_options(options);
    opt.device_id = -2;  // offline: code generation only
    ExprCopy pred;
    if (fused_predicate && fused_predicate->n_nodes > 0) pred = copy_expr(fused_predicate);
    std::vector<std::string> names((size_t)n_group_by);
    AggOp op(copy_aggs(aggs, n_aggs), copy_exprs(group_by, n_group_by), names, n_group_by == 0, pred, opt);
    std::string gen = op.debug_source(cols_of_schema(input_schema));
    if (compile) jit_compile_to_cubin("agg_table+agg", gen, nullptr);
    if (source_out) *source_out = dup_string(jit_full_source("agg_table+agg", gen));
  });
}

int sqlrs_debug_compile_joinagg(const sqlrs_agg_desc* aggs, int32_t n_aggs, const sqlrs_expr* group_by, int32_t n_group_by,
                                const sqlrs_expr* right_keys, int32_t n_keys, const sqlrs_expr* probe_predicate, const sqlrs_expr* join_filter,
                                const ArrowSchema* build_schema, const ArrowSchema* probe_schema, const sqlrs_options* options, int32_t compile,
                                char** source_out) {
  return guarded([&] {
    Options opt = copy_options(options);
    opt.device_id = -2;
    std::vector<std::string> names((size_t)n_group_by);
    AggOp op(copy_aggs(aggs, n_aggs), copy_exprs(group_by, n_group_by), names, n_group_by == 0, {}, opt);
    ExprCopy pp, jf;
    if (probe_predicate && probe_predicate->n_nodes > 0) pp = copy_expr(probe_predicate);
    if (join_filter && join_filter->n_nodes > 0) jf = copy_expr(join_filter);
    std::string gen = op.debug_join_source(cols_of_schema(build_schema), cols_of_schema(probe_schema), copy_exprs(right_keys, n_keys), pp, jf);
    if (compile) jit_compile_to_cubin("agg_table+join_table+joinagg", gen, nullptr);
    if (source_out) *source_out = dup_string(jit_full_source("agg_table+join_table+joinagg", gen));
  });
}

int sqlrs_debug_compile_joinprobe(const sqlrs_expr* right_keys, int32_t n_keys, const sqlrs_expr* probe_predicate, const ArrowSchema* probe_schema,
                                  const sqlrs_options* options, int32_t compile, char** source_out) {
  return guarded([&] {
    Options opt = copy_options(options);
    opt.device_id = -2;
    ExprCopy pp;
    if (probe_predicate && probe_predicate->n_nodes > 0) pp = copy_expr(probe_predicate);
    JoinOp op(SQLRS_JOIN_INNER, copy_exprs(right_keys, n_keys), copy_exprs(right_keys, n_keys), {}, {}, opt);
    std::string gen = op.debug_probe_source(cols_of_schema(probe_schema), pp);
    if (compile) jit_compile_to_cubin("join_table+joinprobe", gen, nullptr);
    // the fused BUILD kernel (csrc/jit/joinbuild.cuh) takes the same generated program for the build side's Filter + key
    if (compile && n_keys == 1 && opt.match_mode == SQLRS_MATCH_HASH_AND_KEY) jit_compile_to_cubin("join_table+joinbuild", gen, nullptr);
    if (source_out) *source_out = dup_string(jit_full_source("join_table+joinprobe", gen));
  });
}

int sqlrs_debug_compile_joinchain(const sqlrs_expr* right_keys1, int32_t n_keys, const sqlrs_expr* probe_predicate, const sqlrs_expr* chain_key,
                                  const ArrowSchema* build_schema, const ArrowSchema* probe_schema, const sqlrs_options* options, int32_t compile,
                                  char** source_out) {
  return guarded([&] {
    Options opt = copy_options(options);
    opt.device_id = -2;
    ExprCopy pp;
    if (probe_predicate && probe_predicate->n_nodes > 0) pp = copy_expr(probe_predicate);
    if (!chain_key) fail(SQLRS_ERR_INVALID_ARG, "chain_key is NULL");
    JoinChainOp op(opt);
    std::string gen = op.debug_source(cols_of_schema(build_schema), cols_of_schema(probe_schema), copy_exprs(right_keys1, n_keys), pp, copy_expr(chain_key));
    if (compile) jit_compile_to_cubin("join_table+joinchain", gen, nullptr);
    if (source_out) *source_out = dup_string(jit_full_source("join_table+joinchain", gen));
  });
}

int sqlrs_debug_compile_eval(const sqlrs_expr* exprs, int32_t n_exprs, int32_t as_keep_mask, const ArrowSchema* input_schema,
                             int32_t compile, char** source_out) {
  return guarded([&] {
    EvalRequest req;
    for (int32_t k = 0; k < n_exprs; k++) {
      req.exprs.push_back(copy_expr(&exprs[k]));
      req.is_key.push_back(false);
      req.outs.push_back({as_keep_mask ? OUT_KEEP : OUT_VALUE, k});
    }
    EvalProgram prog(std::move(req));
    std::vector<int> od, ed;
    std::vector<bool> on;
    std::string gen = prog.source_for(cols_of_schema(input_schema), &od, &on, &ed);
    if (compile) jit_compile_to_cubin("eval", gen, nullptr);
    if (source_out) *source_out = dup_string(jit_full_source("eval", gen));
  });
}

void sqlrs_free(void* p) { std::free(p); }

}  // extern "C"
