/*
 * sqlrs_b200 — C ABI of the B200-native execution backend for Fedomn/sqlrs.
 *
 * This header is the drop-in boundary.  Every entry point is
 * `extern "C"`, takes plain pointers / sizes / Arrow C-Data structs, and
 * replaces one piece of the reference's Rust operator interface.  The
 * reference has no FFI of its own (it is pure Rust), so each function cites the
 * Rust item a maintainer's `arrow::ffi` shim would forward to it
 * (INTEGRATION.md shows that shim).
 *
 * The same ABI is compiled twice:
 *   - libsqlrs_b200.so      (sqlrs_b200/csrc, CUDA sm_100a)  — symbols sqlrs_*
 *   - liboracle.so          (oracle/, CPU restatement, test infrastructure only)
 *                                                            — symbols sqlrs_oracle_*
 * so that the parity tests drive both through identical calls.
 *
 * Ownership rules (Arrow C Data Interface):
 *   - input `ArrowArray` / `ArrowDeviceArray` structs are MOVED into the callee
 *     (it copies the struct, sets `release = NULL` in the caller's copy and
 *     calls the original release callback once it no longer needs the buffers —
 *     possibly after the call returns, when an async H2D copy is still in
 *     flight);
 *   - input `ArrowSchema` structs are only BORROWED for the duration of the call;
 *   - output structs are filled by the callee and carry its release callbacks.
 *
 * Error convention (reference: `Result<RecordBatch, ExecutorError>`,
 * src/executor/mod.rs:67-85): every function returns an `int` status, 0 = ok;
 * `sqlrs_last_error()` returns the thread-local message.  The library never
 * aborts the process; places where the reference panics (`todo!()`,
 * `unimplemented!()`, `expect`) return SQLRS_ERR_UNSUPPORTED / SQLRS_ERR_INTERNAL
 * so a shim can fall back to the CPU operator.
 *
 * Threading (reference: BoxStream is Send + 'static, polled by one task):
 * handles are single-owner, not thread-safe, movable between threads between
 * calls (no thread-affine CUDA state: the device is set on every call and all
 * work goes to the handle's stream).
 */
#ifndef SQLRS_B200_H
#define SQLRS_B200_H

#include <stdint.h>
#include "arrow_c_data.h"

#ifdef SQLRS_ORACLE_BUILD
#define SQLRS_API(name) sqlrs_oracle_##name
#else
#define SQLRS_API(name) sqlrs_##name
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define SQLRS_ABI_VERSION 2

/* ---- status codes: ExecutorError, src/executor/mod.rs:67-85 ------------------- */
#define SQLRS_OK 0
#define SQLRS_ERR_INTERNAL 1    /* ExecutorError::InternalError(String) */
#define SQLRS_ERR_ARROW 2       /* ExecutorError::Arrow(ArrowError): DivideByZero, schema mismatch ... */
#define SQLRS_ERR_UNSUPPORTED 3 /* the reference panics here (todo!/unimplemented!) or the GPU path lacks the dtype */
#define SQLRS_ERR_INVALID_ARG 4
#define SQLRS_ERR_CUDA 5
#define SQLRS_ERR_STORAGE 6     /* ExecutorError::Storage(StorageError): io error / table not found (src/storage/mod.rs) */

/* ---- the type universe of the v1 executor: ScalarValue, src/types/mod.rs:23-36 - */
#define SQLRS_DT_NULL 0
#define SQLRS_DT_BOOL 1
#define SQLRS_DT_INT32 2
#define SQLRS_DT_INT64 3
#define SQLRS_DT_FLOAT64 4
#define SQLRS_DT_UTF8 5 /* CUDA library: dictionary-encoded on ingest (string pool ids in HBM), decoded on export; =, <> compare ids, <, <=, >, >=
                           the pool's byte-wise ranks; MIN / MAX over Utf8 EXPRESSIONS and Utf8 casts return SQLRS_ERR_UNSUPPORTED */

/* ---- expression bytecode: flattened BoundExpr, src/binder/expression/mod.rs:18-27
 * Postfix order: children first (left then right), then the node.  `Alias` is
 * dropped by the host when flattening (evaluator.rs:25 evaluates the inner expr).
 * Operand types must already match — the binder inserts TypeCast nodes
 * (src/binder/expression/binary_op.rs:27-56). */
#define SQLRS_OP_INPUT_REF 1 /* BoundExpr::InputRef{index, return_type}   evaluator.rs:15 */
#define SQLRS_OP_CONSTANT 2  /* BoundExpr::Constant(ScalarValue)           evaluator.rs:21 */
#define SQLRS_OP_CAST 3      /* BoundExpr::TypeCast{expr, cast_type}       evaluator.rs:23 */
#define SQLRS_OP_ADD 10      /* BinaryOperator::Plus      array_compute.rs:76 (wrapping ints) */
#define SQLRS_OP_SUB 11      /* BinaryOperator::Minus     array_compute.rs:77 */
#define SQLRS_OP_MUL 12      /* BinaryOperator::Multiply  array_compute.rs:78 */
#define SQLRS_OP_DIV 13      /* BinaryOperator::Divide    array_compute.rs:79 (DivideByZero -> SQLRS_ERR_ARROW) */
/* the v2 engine's arithmetic (src/function/scalar/arithmetic_function.rs:63-264: add_checked / subtract_checked /
 * multiply_checked / divide_checked): as above, but an Int32 / Int64 result that does not fit its type ends the stream with
 * SQLRS_ERR_ARROW ("Overflow happened on ...") instead of wrapping; Float64 is unchanged */
#define SQLRS_OP_ADD_CHECKED 14
#define SQLRS_OP_SUB_CHECKED 15
#define SQLRS_OP_MUL_CHECKED 16
#define SQLRS_OP_DIV_CHECKED 17
#define SQLRS_OP_GT 20       /* gt_dyn     array_compute.rs:80 */
#define SQLRS_OP_LT 21       /* lt_dyn     array_compute.rs:81 */
#define SQLRS_OP_GE 22       /* gt_eq_dyn  array_compute.rs:82 */
#define SQLRS_OP_LE 23       /* lt_eq_dyn  array_compute.rs:83 */
#define SQLRS_OP_EQ 24       /* eq_dyn     array_compute.rs:84 */
#define SQLRS_OP_NE 25       /* neq_dyn    array_compute.rs:85 */
#define SQLRS_OP_AND 30      /* and_kleene array_compute.rs:86 (Boolean only, else InternalError) */
#define SQLRS_OP_OR 31       /* or_kleene  array_compute.rs:87 */

typedef struct sqlrs_expr_node {
  int32_t op;       /* SQLRS_OP_* */
  int32_t dtype;    /* result type: INPUT_REF -> column type, CONSTANT -> scalar type,
                       CAST -> cast_type, binary op -> return_type */
  int32_t index;    /* INPUT_REF: column index into the input batch */
  int32_t is_null;  /* CONSTANT: 1 = ScalarValue::X(None) */
  int64_t imm_bits; /* CONSTANT: bool / i32 / i64 value, or the IEEE-754 bits of an f64 */
  const char* str;  /* CONSTANT of SQLRS_DT_UTF8: NUL-terminated UTF-8, else NULL */
} sqlrs_expr_node;

typedef struct sqlrs_expr {
  const sqlrs_expr_node* nodes;
  int32_t n_nodes; /* 0 = "no expression" where an expression is optional */
} sqlrs_expr;

/* ---- aggregates: BoundAggFunc, src/binder/expression/agg_func.rs:10-34 --------- */
#define SQLRS_AGG_COUNT 0 /* CountAccumulator  aggregate/count.rs:10-29 */
#define SQLRS_AGG_SUM 1   /* SumAccumulator    aggregate/sum.rs:36-97 */
#define SQLRS_AGG_MIN 2   /* MinAccumulator    aggregate/min_max.rs:111-133 */
#define SQLRS_AGG_MAX 3   /* MaxAccumulator    aggregate/min_max.rs:135-157 */

typedef struct sqlrs_agg_desc {
  int32_t func;         /* SQLRS_AGG_* */
  int32_t distinct;     /* BoundAggFunc::distinct — Count/Sum keep a set of values per group (count.rs:31-58, sum.rs:99-132) */
  int32_t return_dtype; /* BoundAggFunc::return_type */
  int32_t reserved;
  sqlrs_expr arg;       /* exprs[0] — only the first argument is evaluated (hash_agg.rs:65) */
  const char* name;     /* output field name, e.g. "Sum(b)" (evaluator.rs:52-56); computed by the host */
} sqlrs_agg_desc;

/* ---- JoinType, src/binder/table/join.rs:17-24 ---------------------------------- */
#define SQLRS_JOIN_INNER 0
#define SQLRS_JOIN_LEFT 1
#define SQLRS_JOIN_RIGHT 2
#define SQLRS_JOIN_FULL 3

/* ---- behaviour switches for the reference quirks K1/K2 ----- */
#define SQLRS_COUNT_REFERENCE_OVERWRITE 0 /* count.rs:22 assigns: value = count in the LAST batch touching the group */
#define SQLRS_COUNT_SQL_ACCUMULATE 1      /* += (SQL semantics) */
#define SQLRS_MATCH_HASH_ONLY 0           /* hash_agg.rs:87-110 / hash_join.rs:222-233: 64-bit row hash is the identity */
#define SQLRS_MATCH_HASH_AND_KEY 1        /* also compare key values; NULL keys never join (SQL semantics) */

#define SQLRS_FLAG_NO_FUSION 1   /* plan API: never pick a fused pipeline, run operator by operator */
#define SQLRS_FLAG_DEVICE_OUTPUT 2 /* reserved */
#define SQLRS_FLAG_TIMING 4 /* bracket the dominant scan kernels with CUDA events (sqlrs_plan_scan_kernel_ms) */
#define SQLRS_FLAG_KERNEL_EVENTS 8 /* record CUDA events around every hot kernel; adds no synchronisation (sqlrs_kernel_events_collect) */

typedef struct sqlrs_options {
  int32_t count_mode; /* default SQLRS_COUNT_REFERENCE_OVERWRITE */
  int32_t match_mode; /* default SQLRS_MATCH_HASH_ONLY */
  int32_t device_id;  /* CUDA ordinal, -1 = current device (ignored by the oracle) */
  int32_t flags;      /* SQLRS_FLAG_* */
  void* stream;       /* cudaStream_t to launch on; NULL = library-owned non-blocking stream */
} sqlrs_options;

/* ---- library -------------------------------------------------------------------- */
int SQLRS_API(abi_version)(void);
const char* SQLRS_API(last_error)(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t SQLRS_API(kernel_launches)(void);

/* row hashes exactly as create_hashes (src/executor/aggregate/hash_utils.rs:161-220)
 * with RandomState::with_seeds(0,0,0,0): `columns` is a struct array whose children
 * are the key columns; writes `length` u64 into out_hashes (host memory). */
int SQLRS_API(create_hashes)(struct ArrowArray* columns, const struct ArrowSchema* schema,
                             uint64_t* out_hashes);

/* BoundExpr::eval_column (src/executor/evaluator.rs:13-28): evaluates one expression
 * over a batch, returns a 1-column batch. */
int SQLRS_API(eval_expr)(const sqlrs_expr* expr, const sqlrs_options* options,
                         struct ArrowArray* batch, const struct ArrowSchema* schema,
                         struct ArrowArray* out, struct ArrowSchema* out_schema);

/* ---- FilterExecutor{expr, child}, src/executor/filter.rs:7-26 --------------------
 * one output batch per input batch, row order preserved, NULL predicate drops the row */
typedef struct sqlrs_filter sqlrs_filter;
int SQLRS_API(filter_create)(const sqlrs_expr* predicate, const sqlrs_options* options,
                             sqlrs_filter** out);
int SQLRS_API(filter_execute)(sqlrs_filter* f, struct ArrowArray* batch,
                              const struct ArrowSchema* schema, struct ArrowArray* out,
                              struct ArrowSchema* out_schema);
void SQLRS_API(filter_destroy)(sqlrs_filter* f);

/* ---- SimpleAggExecutor{agg_funcs, child}, src/executor/aggregate/simple_agg.rs:10-65
 * push every child batch in order, then finish -> exactly one 1-row batch.
 * finish with zero pushed batches is an error (simple_agg.rs:63 unwraps None). */
typedef struct sqlrs_simple_agg sqlrs_simple_agg;
int SQLRS_API(simple_agg_create)(const sqlrs_agg_desc* aggs, int32_t n_aggs,
                                 const sqlrs_options* options, sqlrs_simple_agg** out);
int SQLRS_API(simple_agg_push)(sqlrs_simple_agg* a, struct ArrowArray* batch,
                               const struct ArrowSchema* schema);
int SQLRS_API(simple_agg_finish)(sqlrs_simple_agg* a, struct ArrowArray* out,
                                 struct ArrowSchema* out_schema);
void SQLRS_API(simple_agg_destroy)(sqlrs_simple_agg* a);

/* ---- HashAggExecutor{agg_funcs, group_by, child}, src/executor/aggregate/hash_agg.rs:15-150
 * output: one batch, columns = group keys then aggregates, rows in first-appearance
 * order of the group (hash_agg.rs:98,134).  finish with zero pushed batches is an
 * error (hash_agg.rs:125 unwraps None). */
typedef struct sqlrs_hash_agg sqlrs_hash_agg;
int SQLRS_API(hash_agg_create)(const sqlrs_agg_desc* aggs, int32_t n_aggs,
                               const sqlrs_expr* group_by, const char* const* group_names,
                               int32_t n_group_by, const sqlrs_options* options,
                               sqlrs_hash_agg** out);
int SQLRS_API(hash_agg_push)(sqlrs_hash_agg* a, struct ArrowArray* batch,
                             const struct ArrowSchema* schema);
int SQLRS_API(hash_agg_finish)(sqlrs_hash_agg* a, struct ArrowArray* out,
                               struct ArrowSchema* out_schema);
void SQLRS_API(hash_agg_destroy)(sqlrs_hash_agg* a);

/* ---- HashJoinExecutor{left_child, right_child, join_type, join_condition,
 *      join_output_schema}, src/executor/join/hash_join.rs:16-23,147-323 -----------
 * left child = build side: push all its batches (build_push), then stream the right
 * child through probe (one output batch per probe batch, hash_join.rs:208-292), then
 * finish yields the Left/Full tail of unmatched build rows (hash_join.rs:296-322).
 * *has_batch = 0 where the reference yields nothing (empty build side :183-185,
 * Inner/Right tail).  `join_output_schema` = the struct schema of the joined row
 * (all left fields then all right fields, names "table.col", catalog/mod.rs:131-138). */
typedef struct sqlrs_hash_join sqlrs_hash_join;
int SQLRS_API(hash_join_create)(int32_t join_type, const sqlrs_expr* left_keys,
                                const sqlrs_expr* right_keys, int32_t n_keys,
                                const sqlrs_expr* filter, /* NULL or n_nodes==0: no non-equi filter */
                                const struct ArrowSchema* join_output_schema,
                                const sqlrs_options* options, sqlrs_hash_join** out);
int SQLRS_API(hash_join_build_push)(sqlrs_hash_join* j, struct ArrowArray* batch,
                                    const struct ArrowSchema* schema);
int SQLRS_API(hash_join_probe)(sqlrs_hash_join* j, struct ArrowArray* batch,
                               const struct ArrowSchema* schema, struct ArrowArray* out,
                               struct ArrowSchema* out_schema, int32_t* has_batch);
int SQLRS_API(hash_join_finish)(sqlrs_hash_join* j, struct ArrowArray* out,
                                struct ArrowSchema* out_schema, int32_t* has_batch);
void SQLRS_API(hash_join_destroy)(sqlrs_hash_join* j);

/* ---- CrossJoinExecutor{left_child, right_child, join_output_schema}, src/executor/join/cross_join.rs:8-57 -----------
 * (what comma joins and `cross join` plan to).  Push every left batch (build_push), then each
 * right batch through probe: it queues ONE output batch per left row — that row's values repeated to the right batch's
 * length, then the right batch's columns (cross_join.rs:41-55) — which cross_join_next hands out in order
 * (*has_batch = 0 when the queue is empty).  An empty left side yields nothing (:33-35). */
typedef struct sqlrs_cross_join sqlrs_cross_join;
int SQLRS_API(cross_join_create)(const struct ArrowSchema* join_output_schema, const sqlrs_options* options,
                                 sqlrs_cross_join** out);
int SQLRS_API(cross_join_build_push)(sqlrs_cross_join* j, struct ArrowArray* batch, const struct ArrowSchema* schema);
int SQLRS_API(cross_join_probe)(sqlrs_cross_join* j, struct ArrowArray* batch, const struct ArrowSchema* schema);
int SQLRS_API(cross_join_next)(sqlrs_cross_join* j, struct ArrowArray* out, struct ArrowSchema* out_schema,
                               int32_t* has_batch);
void SQLRS_API(cross_join_destroy)(sqlrs_cross_join* j);

/* ==== the operators that FOLLOW the hot path in a v1 plan; the planner
 *      stacks them Agg -> Order -> Project -> Limit (src/planner/select.rs:34-45) ================= */

/* ---- ProjectExecutor{exprs, child}, src/executor/project.rs:6-29 ------------------
 * one output batch per input batch, columns = eval_column of each expression.  Output fields as
 * eval_field (evaluator.rs:30-64): names[k] == NULL (only for a bare InputRef) keeps the input field,
 * name and nullability; otherwise the field is (names[k], result type, nullable) — the host computes
 * the name (binary-op / cast / alias naming rules). */
typedef struct sqlrs_project sqlrs_project;
int SQLRS_API(project_create)(const sqlrs_expr* exprs, const char* const* names, int32_t n_exprs,
                              const sqlrs_options* options, sqlrs_project** out);
int SQLRS_API(project_execute)(sqlrs_project* p, struct ArrowArray* batch,
                               const struct ArrowSchema* schema, struct ArrowArray* out,
                               struct ArrowSchema* out_schema);
void SQLRS_API(project_destroy)(sqlrs_project* p);

/* ---- OrderExecutor{order_by, child}, src/executor/order.rs:8-67 -------------------
 * push every child batch, then finish -> ONE batch: all rows (concat_batches :28) taken in the order of
 * lexsort_to_indices over the evaluated sort expressions (:30-45).  asc[k] = BoundOrderBy::asc;
 * NULLs sort first whatever the direction (SortOptions::default().nulls_first).  Rows equal on every key:
 * the reference sorts unstably (order undefined); this ABI keeps them in input order.  finish with zero
 * pushed batches is an error (order.rs:27 unwraps None). */
typedef struct sqlrs_order sqlrs_order;
int SQLRS_API(order_create)(const sqlrs_expr* order_by, const int32_t* asc, int32_t n_order_by,
                            const sqlrs_options* options, sqlrs_order** out);
int SQLRS_API(order_push)(sqlrs_order* o, struct ArrowArray* batch, const struct ArrowSchema* schema);
int SQLRS_API(order_finish)(sqlrs_order* o, struct ArrowArray* out, struct ArrowSchema* out_schema);
void SQLRS_API(order_destroy)(sqlrs_order* o);

/* ---- LimitExecutor{limit, offset, child}, src/executor/limit.rs:6-80 ---------------
 * limit / offset = the bound constants, -1 = None.  Push the child's batches in order: *has_batch = 1
 * when the call yields an output batch (the input itself or a slice of it), *done = 1 once the reference
 * would stop pulling from its child (limit.rs:31-33 limit 0, :76-78 break).  Reproduces the reference's
 * per-batch arithmetic, including `limit: None` meaning "the current batch's row count" (:40). */
typedef struct sqlrs_limit sqlrs_limit;
int SQLRS_API(limit_create)(int64_t limit, int64_t offset, const sqlrs_options* options, sqlrs_limit** out);
int SQLRS_API(limit_push)(sqlrs_limit* l, struct ArrowArray* batch, const struct ArrowSchema* schema,
                          struct ArrowArray* out, struct ArrowSchema* out_schema, int32_t* has_batch,
                          int32_t* done);
void SQLRS_API(limit_destroy)(sqlrs_limit* l);

/* ---- whole physical sub-plan: what ExecutorBuilder::build(plan) (src/executor/mod.rs:45-47,
 *      visit_* :87-200) wires together.  Handing the GPU the subtree instead of one
 *      operator lets it keep intermediates in HBM and pick a fused pipeline
 *      (scan->filter->agg, scan->filter->join->join->agg) when the shape is registered;
 *      otherwise it composes the per-operator kernels above.  Results are identical
 *      either way. ------------------------------------------------------------------ */
#define SQLRS_NODE_SCAN 1       /* PhysicalTableScan: batches pushed by the host per table_slot */
#define SQLRS_NODE_FILTER 2     /* PhysicalFilter     mod.rs:139-149 */
#define SQLRS_NODE_SIMPLE_AGG 3 /* PhysicalSimpleAgg  mod.rs:151-161 */
#define SQLRS_NODE_HASH_AGG 4   /* PhysicalHashAgg    mod.rs:163-174 */
#define SQLRS_NODE_HASH_JOIN 5  /* PhysicalHashJoin   mod.rs:103-114 (child0 = left = build) */
#define SQLRS_NODE_PROJECT 6    /* PhysicalProject    mod.rs:127-137 */
#define SQLRS_NODE_ORDER 7      /* PhysicalOrder      mod.rs:189-199 */
#define SQLRS_NODE_LIMIT 8      /* PhysicalLimit      mod.rs:176-187 */
#define SQLRS_NODE_CROSS_JOIN 9 /* PhysicalCrossJoin  mod.rs:116-125 (child0 = left, child1 = right; join_output_schema) */

typedef struct sqlrs_plan_node {
  int32_t kind;   /* SQLRS_NODE_* */
  int32_t child0; /* index into the node array, -1 = none */
  int32_t child1; /* HASH_JOIN: right (probe) child */
  int32_t table_slot; /* SCAN */
  sqlrs_expr predicate; /* FILTER; HASH_JOIN: optional non-equi filter */
  const sqlrs_agg_desc* aggs; /* SIMPLE_AGG / HASH_AGG */
  int32_t n_aggs;
  int32_t n_group_by;
  const sqlrs_expr* group_by; /* HASH_AGG */
  const char* const* group_names;
  int32_t join_type; /* HASH_JOIN */
  int32_t n_keys;
  const sqlrs_expr* left_keys;
  const sqlrs_expr* right_keys;
  const struct ArrowSchema* join_output_schema;
  /* ABI version 2 */
  const sqlrs_expr* exprs;        /* PROJECT: select list; ORDER: sort expressions */
  const char* const* expr_names;  /* PROJECT: output field names (host-computed eval_field names) */
  const int32_t* order_asc;       /* ORDER: BoundOrderBy::asc per sort expression */
  int32_t n_exprs;
  int32_t reserved;
  int64_t limit;  /* LIMIT: -1 = None */
  int64_t offset; /* LIMIT: -1 = None */
} sqlrs_plan_node;

typedef struct sqlrs_plan sqlrs_plan;
int SQLRS_API(plan_create)(const sqlrs_plan_node* nodes, int32_t n_nodes, int32_t root,
                           const sqlrs_options* options, sqlrs_plan** out);
/* one call per batch of the table bound to `table_slot`, in stream order */
int SQLRS_API(plan_push_table)(sqlrs_plan* p, int32_t table_slot, struct ArrowArray* batch,
                               const struct ArrowSchema* schema);
/* same with buffers already resident in HBM (zero copy; device pointers in buffers[]) */
int SQLRS_API(plan_push_table_device)(sqlrs_plan* p, int32_t table_slot,
                                      struct ArrowDeviceArray* batch,
                                      const struct ArrowSchema* schema);
/* runs the plan over everything pushed so far; work is enqueued on the plan's stream
 * and complete when the call returns only for host-visible results (plan_next syncs) */
int SQLRS_API(plan_execute)(sqlrs_plan* p);
/* pull the result stream (try_collect, src/executor/mod.rs:58-64) */
/* The next pending result batch WITHOUT leaving the device (the exchange steps of multi-GPU execution): its shape, and a copy
 * of its columns into caller-provided DEVICE buffers (columns[c] holds >= n_rows values of the column's width; the copy is
 * enqueued on the plan's stream).  Columns with NULLs are refused with SQLRS_ERR_UNSUPPORTED (take sqlrs_plan_next).
 * *has_batch = 0 when nothing is pending. */
int SQLRS_API(plan_result_shape)(sqlrs_plan* p, int64_t* n_rows, int32_t* n_columns, int32_t* has_batch);
int SQLRS_API(plan_next_to_device)(sqlrs_plan* p, void* const* columns, int32_t n_columns);
/* sqlrs_plan_push_table of ONE host batch that the plan then scans in slices of batch_rows rows (a multiple of 32), the
 * way the reference's scan hands its executor 1024-row batches (src/storage/csv.rs:105): one H2D copy, zero-copy slices. */
int SQLRS_API(plan_push_table_batched)(sqlrs_plan* p, int32_t table_slot, struct ArrowArray* batch,
                                       const struct ArrowSchema* schema, int64_t batch_rows);
int SQLRS_API(plan_next)(sqlrs_plan* p, struct ArrowArray* out, struct ArrowSchema* out_schema,
                         int32_t* has_batch);
/* forget pushed tables / operator state, keep plan + device scratch (repeated runs) */
int SQLRS_API(plan_reset)(sqlrs_plan* p);
/* forget the batches pushed into ONE table slot (the other slots keep theirs: a plan re-run with one input replaced) */
int SQLRS_API(plan_clear_table)(sqlrs_plan* p, int32_t table_slot);
/* human-readable: which pipeline (fused / generic) and kernels the plan runs with */
const char* SQLRS_API(plan_describe)(sqlrs_plan* p);
void SQLRS_API(plan_destroy)(sqlrs_plan* p);
/* ---- GPU-resident tables: the InMemoryStorage / InMemoryTable analogue (src/storage/memory.rs:38-47,137-170;
 *      SURVEY.md §8f rank 2).  `create_mem_table(id, Vec<RecordBatch>)` becomes table_create + one table_append per
 *      batch: the batch is moved in, copied to HBM once and kept there; any number of plans then scan it with
 *      plan_push_table_resident without crossing PCIe again (zero copy; the table may be destroyed while plans still
 *      hold its batches).  table_read mirrors InMemoryTransaction::next_batch (memory.rs:151-170): batch `batch_index`
 *      back on the host, *has_batch = 0 past the end; `projection` (column indices, NULL = all) is honoured here although
 *      the reference's in-memory table ignores it (:137-143). ------------------------------------------------------- */
typedef struct sqlrs_table sqlrs_table;
int SQLRS_API(table_create)(const sqlrs_options* options, sqlrs_table** out);
int SQLRS_API(table_append)(sqlrs_table* t, struct ArrowArray* batch, const struct ArrowSchema* schema);
int64_t SQLRS_API(table_num_rows)(sqlrs_table* t);
int32_t SQLRS_API(table_num_batches)(sqlrs_table* t);
int SQLRS_API(table_read)(sqlrs_table* t, int32_t batch_index, const int32_t* projection, int32_t n_projection,
                          struct ArrowArray* out, struct ArrowSchema* out_schema, int32_t* has_batch);
void SQLRS_API(table_destroy)(sqlrs_table* t);
/* every batch of the table, in order, as the batches of `table_slot` (what PhysicalTableScan pulls, table_scan.rs:16-35) */
/* CsvTable (src/storage/csv.rs:99-235): the file parsed ON THE DEVICE into a resident table of batch_rows-row batches (1024 in
 * the reference, :105; a multiple of 32 here).  has_header / delimiter as CsvConfig (:99-109); the schema is inferred from the first
 * 10 records like arrow-csv's (Boolean / Int64 / Float64 / Utf8, every field nullable); bounds_offset / bounds_limit < 0 = no
 * bounds, else Bounds = (offset, limit) over the whole file (:196-206); projection == NULL = every column.  Scan it with
 * sqlrs_plan_push_table_resident, read it back with sqlrs_table_read. */
int SQLRS_API(table_read_csv)(const char* path, int32_t has_header, int32_t delimiter, int64_t batch_rows, int64_t bounds_offset,
                              int64_t bounds_limit, const int32_t* projection, int32_t n_projection, const sqlrs_options* options,
                              sqlrs_table** out);
int SQLRS_API(plan_push_table_resident)(sqlrs_plan* p, int32_t table_slot, sqlrs_table* t);

/* ---- partial / final aggregation for plans sharded over several GPUs.  The plan
 *      root must be an aggregate.  Each rank: push its shard, execute_partial (row_base = global row
 *      id of the shard's first row, keeps first-appearance order global), export_partials -> a host
 *      batch of un-finalised groups whose column 0 ("hash", int64) is the group identity to
 *      radix-partition on, the rest is opaque to the host; after the exchange: clear_partials,
 *      merge_partials for every received batch (its own share included), finish_partial, plan_next. */
int SQLRS_API(plan_execute_partial)(sqlrs_plan* p, int64_t row_base);
int SQLRS_API(plan_export_partials)(sqlrs_plan* p, struct ArrowArray* out, struct ArrowSchema* out_schema);
int SQLRS_API(plan_clear_partials)(sqlrs_plan* p);
int SQLRS_API(plan_merge_partials)(sqlrs_plan* p, struct ArrowArray* partials, const struct ArrowSchema* schema);
int SQLRS_API(plan_finish_partial)(sqlrs_plan* p);
/* DISTINCT aggregates (reference: a HashSet<ScalarValue> per group, count.rs:31-58 / sum.rs:99-132) keep their set elements in
 * one more table per DISTINCT aggregate, next to the table of the plain aggregates.  *n_tables = how many tables the partial
 * state of this plan has (1 without DISTINCT); select_partials_table makes export / clear / merge / row_words (and the
 * _device forms below) address table `index` (execute_partial selects 0).  Exchange every table the same way, merge table 0
 * first, then finish_partial once.  Column 0 of every table's rows is the hash to partition on. */
int SQLRS_API(plan_partials_tables)(sqlrs_plan* p, int32_t* n_tables);
int SQLRS_API(plan_select_partials_table)(sqlrs_plan* p, int32_t index);
/* the same exchange without leaving HBM (what the NCCL path uses): partial groups packed row-major into
 * caller-provided DEVICE memory on the plan's stream — (cap_rows + 1) rows of *n_words u64 each, row 0 =
 * header {number of groups (may exceed cap_rows: then rows are missing and the caller must fall back to the
 * host path)}, row 1+i = [hash, min_row, null mask, key bits..., accumulator words...] — and the merge of
 * n_buffers such buffers laid out back to back (e.g. the output of an all-gather).  The oracle build returns
 * SQLRS_ERR_UNSUPPORTED for the two _device calls.
 * Stream contract: with sqlrs_options.stream set, export writes `dst` and merge reads `src` ON THAT STREAM — enqueue the
 * collective on the same stream (or order it with events).  With options.stream == NULL the plan owns a private
 * non-blocking stream: export then synchronises before it returns (dst is complete for any stream), and the caller
 * must have completed the writes to `src` (synchronise its own stream) before calling merge. */
int SQLRS_API(plan_partials_row_words)(sqlrs_plan* p, int32_t* n_words);
/* Radix partitioning of the partial groups for the all-to-all exchange ("radix-partition partial rows by
 * hash(keys) mod G"): the groups of the un-finalised table are packed into `n_parts` back-to-back regions of `dst` (DEVICE
 * memory, on the plan's stream), region q = (cap_rows + 1) rows of *n_words u64 in the layout above, holding the groups
 * whose identity hash (as u64) mod n_parts == q; its row 0 = header {count (may exceed cap_rows: rows missing, retry larger)}.
 * Equal-sized regions, so ONE all_to_all_single with equal splits moves them and the receive buffer feeds
 * sqlrs_plan_merge_partials_device(n_buffers = n_parts) directly.  *groups_out = number of groups in the table (this call
 * synchronises once to learn it; pass cap_rows <= 0 to only learn the count). */
int SQLRS_API(plan_export_partials_partitioned)(sqlrs_plan* p, void* dst, int32_t n_parts, int64_t cap_rows, int64_t* groups_out);
int SQLRS_API(plan_export_partials_device)(sqlrs_plan* p, void* dst, int64_t cap_rows);
int SQLRS_API(plan_merge_partials_device)(sqlrs_plan* p, const void* src, int32_t n_buffers, int64_t cap_rows);
/* SQLRS_FLAG_TIMING: device time (ms, CUDA events on the plan's stream) the dominant scan kernel(s) of the
 * last plan_execute took, and how many launches that covers — bench.py's roofline numerator/denominator */
double SQLRS_API(plan_scan_kernel_ms)(sqlrs_plan* p, int64_t* n_launches);

/* ---- synthetic TPC-H-shaped tables: counter-based, identical in the
 *      CUDA generator (writes straight into HBM) and the oracle's CPU generator ------ */
#define SQLRS_TPCH_CUSTOMER 0 /* c_custkey i64, c_mktsegment i64 */
#define SQLRS_TPCH_ORDERS 1   /* o_orderkey, o_custkey, o_orderdate, o_shippriority (all i64) */
#define SQLRS_TPCH_LINEITEM 2 /* l_orderkey i64, l_quantity f64, l_extendedprice f64, l_discount f64,
                                 l_tax f64, l_returnflag i64, l_linestatus i64, l_shipdate i64,
                                 l_quantity_i64 i64 */
#define SQLRS_TPCH_FLAGS_8GROUP 0 /* returnflag = u mod 4, linestatus = (u>>2) mod 2 */
#define SQLRS_TPCH_FLAGS_SPEC 1   /* derived from dates as in the TPC-H spec (4 populated groups) */

typedef struct sqlrs_tpch_dims {
  int64_t n_customer; /* 150 000 x SF */
  int64_t n_orders;   /* 1 500 000 x SF */
  int32_t flags_mode; /* SQLRS_TPCH_FLAGS_* */
  int32_t reserved;
} sqlrs_tpch_dims;

int32_t SQLRS_API(tpch_num_columns)(int32_t table);
int64_t SQLRS_API(tpch_num_rows)(const sqlrs_tpch_dims* dims, int32_t table);
/* fills rows [row_begin, row_end) of every column of `table`; columns[c] points to
 * (row_end-row_begin) 8-byte values — device memory for the CUDA library (enqueued on
 * `stream`), host memory for the oracle. */
int SQLRS_API(tpch_generate)(const sqlrs_tpch_dims* dims, int32_t table, int64_t row_begin,
                             int64_t row_end, void* const* columns, void* stream);

/* ---- diagnostics / build check (need no GPU): the CUDA source the library generates for an
 *      operator over batches of `input_schema` (nullable fields are assumed to carry a validity
 *      bitmap), optionally compiled with NVRTC for sm_100a into the on-disk module cache.
 *      *source_out is malloc'ed; release it with sqlrs_free.  The oracle build returns
 *      SQLRS_ERR_UNSUPPORTED. ------------------------------------------------------------------ */
int SQLRS_API(debug_compile_agg)(const sqlrs_agg_desc* aggs, int32_t n_aggs, const sqlrs_expr* group_by,
                                 int32_t n_group_by, const sqlrs_expr* fused_predicate,
                                 const struct ArrowSchema* input_schema, const sqlrs_options* options,
                                 int32_t compile, char** source_out);
/* the fused probe -> aggregate kernel (csrc/jit/joinagg.cuh) for an inner join with `right_keys` over `probe_schema`
 * batches, build side `build_schema`; group_by / aggregate arguments index the joined row (build columns first) */
int SQLRS_API(debug_compile_joinagg)(const sqlrs_agg_desc* aggs, int32_t n_aggs, const sqlrs_expr* group_by,
                                     int32_t n_group_by, const sqlrs_expr* right_keys, int32_t n_keys,
                                     const sqlrs_expr* probe_predicate, const sqlrs_expr* join_filter,
                                     const struct ArrowSchema* build_schema, const struct ArrowSchema* probe_schema,
                                     const sqlrs_options* options, int32_t compile, char** source_out);
/* the fused probe kernel (csrc/jit/joinprobe.cuh) of a hash join with `right_keys` over `probe_schema` batches and an
 * optional Filter fused below the join on the probe side */
int SQLRS_API(debug_compile_joinprobe)(const sqlrs_expr* right_keys, int32_t n_keys, const sqlrs_expr* probe_predicate,
                                       const struct ArrowSchema* probe_schema, const sqlrs_options* options,
                                       int32_t compile, char** source_out);
/* the fused probe -> build kernel of a join chain (csrc/jit/joinchain.cuh): join 1 (`right_keys1` over `probe_schema` batches,
 * build side `build_schema`, optional probe-side Filter) inserts `chain_key` — join 2's build key, an expression over join 1's
 * output row (build columns first) — into join 2's table */
int SQLRS_API(debug_compile_joinchain)(const sqlrs_expr* right_keys1, int32_t n_keys, const sqlrs_expr* probe_predicate,
                                       const sqlrs_expr* chain_key, const struct ArrowSchema* build_schema,
                                       const struct ArrowSchema* probe_schema, const sqlrs_options* options,
                                       int32_t compile, char** source_out);
int SQLRS_API(debug_compile_eval)(const sqlrs_expr* exprs, int32_t n_exprs, int32_t as_keep_mask,
                                  const struct ArrowSchema* input_schema, int32_t compile, char** source_out);
/* SQLRS_FLAG_KERNEL_EVENTS: JSON {"kernel": {"ms": total device time, "launches": n}, ...} of the events recorded since the
 * last call (call after synchronising the stream).  malloc'ed; release with sqlrs_free.  Oracle: "{}". */
int SQLRS_API(kernel_events_collect)(char** json_out);
void SQLRS_API(free)(void* p);

#ifdef __cplusplus
}
#endif

#endif /* SQLRS_B200_H */
