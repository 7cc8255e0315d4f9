"""ctypes binding of include/sqlrs_b200.h — the C-ABI drop-in boundary.

`Library(path, prefix)` binds one build of the ABI.  The product library is
`sqlrs_b200/csrc/libsqlrs_b200.so` (prefix ``sqlrs_``, CUDA sm_100a); `load()` returns it and
raises if it has not been built — there is NO CPU fallback in this package.  The same class can
bind any other build of the ABI by path + symbol prefix (the test-suite binds its CPU checker
that way); nothing in this package does.
"""
from __future__ import annotations

import ctypes as C
import os
import struct

import pyarrow as pa

# ---- constants (mirror include/sqlrs_b200.h) ---------------------------------------------------
OK, ERR_INTERNAL, ERR_ARROW, ERR_UNSUPPORTED, ERR_INVALID_ARG, ERR_CUDA, ERR_STORAGE = range(7)
DT_NULL, DT_BOOL, DT_INT32, DT_INT64, DT_FLOAT64, DT_UTF8 = range(6)
OP_INPUT_REF, OP_CONSTANT, OP_CAST = 1, 2, 3
OP_ADD, OP_SUB, OP_MUL, OP_DIV = 10, 11, 12, 13
OP_ADD_CHECKED, OP_SUB_CHECKED, OP_MUL_CHECKED, OP_DIV_CHECKED = 14, 15, 16, 17
OP_GT, OP_LT, OP_GE, OP_LE, OP_EQ, OP_NE = 20, 21, 22, 23, 24, 25
OP_AND, OP_OR = 30, 31
AGG_COUNT, AGG_SUM, AGG_MIN, AGG_MAX = 0, 1, 2, 3
JOIN_INNER, JOIN_LEFT, JOIN_RIGHT, JOIN_FULL = 0, 1, 2, 3
COUNT_REFERENCE_OVERWRITE, COUNT_SQL_ACCUMULATE = 0, 1
MATCH_HASH_ONLY, MATCH_HASH_AND_KEY = 0, 1
FLAG_NO_FUSION = 1
FLAG_TIMING = 4
FLAG_KERNEL_EVENTS = 8
NODE_SCAN, NODE_FILTER, NODE_SIMPLE_AGG, NODE_HASH_AGG, NODE_HASH_JOIN, NODE_PROJECT, NODE_ORDER, NODE_LIMIT, NODE_CROSS_JOIN = 1, 2, 3, 4, 5, 6, 7, 8, 9
ABI_VERSION = 2
TPCH_CUSTOMER, TPCH_ORDERS, TPCH_LINEITEM = 0, 1, 2
TPCH_FLAGS_8GROUP, TPCH_FLAGS_SPEC = 0, 1

_PA_TO_DT = {
    pa.null(): DT_NULL,
    pa.bool_(): DT_BOOL,
    pa.int32(): DT_INT32,
    pa.int64(): DT_INT64,
    pa.float64(): DT_FLOAT64,
    pa.utf8(): DT_UTF8,
}
_DT_TO_PA = {v: k for k, v in _PA_TO_DT.items()}
# arrow-rs `DataType` Display (used in cast field names, evaluator.rs:44)
DT_NAME = {DT_NULL: "Null", DT_BOOL: "Boolean", DT_INT32: "Int32", DT_INT64: "Int64", DT_FLOAT64: "Float64", DT_UTF8: "Utf8"}


def dtype_of(pa_type) -> int:
    try:
        return _PA_TO_DT[pa_type]
    except KeyError:
        raise ExecutorError(ERR_UNSUPPORTED, f"type {pa_type} is outside the v1 type universe") from None


def pa_type_of(dt: int):
    return _DT_TO_PA[dt]


class ExecutorError(RuntimeError):
    """ExecutorError of src/executor/mod.rs:67-85 (status code + message)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code
        self.message = message


# ---- Arrow C data structs ----------------------------------------------------------------------
class ArrowSchema(C.Structure):
    pass


ArrowSchema._fields_ = [
    ("format", C.c_char_p),
    ("name", C.c_char_p),
    ("metadata", C.c_char_p),
    ("flags", C.c_int64),
    ("n_children", C.c_int64),
    ("children", C.POINTER(C.POINTER(ArrowSchema))),
    ("dictionary", C.POINTER(ArrowSchema)),
    ("release", C.c_void_p),
    ("private_data", C.c_void_p),
]


class ArrowArray(C.Structure):
    pass


ArrowArray._fields_ = [
    ("length", C.c_int64),
    ("null_count", C.c_int64),
    ("offset", C.c_int64),
    ("n_buffers", C.c_int64),
    ("n_children", C.c_int64),
    ("buffers", C.POINTER(C.c_void_p)),
    ("children", C.POINTER(C.POINTER(ArrowArray))),
    ("dictionary", C.POINTER(ArrowArray)),
    ("release", C.c_void_p),
    ("private_data", C.c_void_p),
]


class ArrowDeviceArray(C.Structure):
    _fields_ = [
        ("array", ArrowArray),
        ("device_id", C.c_int64),
        ("device_type", C.c_int32),
        ("sync_event", C.c_void_p),
        ("reserved", C.c_int64 * 3),
    ]


class ExprNode(C.Structure):
    _fields_ = [
        ("op", C.c_int32),
        ("dtype", C.c_int32),
        ("index", C.c_int32),
        ("is_null", C.c_int32),
        ("imm_bits", C.c_int64),
        ("str", C.c_char_p),
    ]


class Expr(C.Structure):
    _fields_ = [("nodes", C.POINTER(ExprNode)), ("n_nodes", C.c_int32)]


class AggDesc(C.Structure):
    _fields_ = [
        ("func", C.c_int32),
        ("distinct", C.c_int32),
        ("return_dtype", C.c_int32),
        ("reserved", C.c_int32),
        ("arg", Expr),
        ("name", C.c_char_p),
    ]


class Options(C.Structure):
    _fields_ = [
        ("count_mode", C.c_int32),
        ("match_mode", C.c_int32),
        ("device_id", C.c_int32),
        ("flags", C.c_int32),
        ("stream", C.c_void_p),
    ]


class PlanNode(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("child0", C.c_int32),
        ("child1", C.c_int32),
        ("table_slot", C.c_int32),
        ("predicate", Expr),
        ("aggs", C.POINTER(AggDesc)),
        ("n_aggs", C.c_int32),
        ("n_group_by", C.c_int32),
        ("group_by", C.POINTER(Expr)),
        ("group_names", C.POINTER(C.c_char_p)),
        ("join_type", C.c_int32),
        ("n_keys", C.c_int32),
        ("left_keys", C.POINTER(Expr)),
        ("right_keys", C.POINTER(Expr)),
        ("join_output_schema", C.POINTER(ArrowSchema)),
        ("exprs", C.POINTER(Expr)),
        ("expr_names", C.POINTER(C.c_char_p)),
        ("order_asc", C.POINTER(C.c_int32)),
        ("n_exprs", C.c_int32),
        ("reserved", C.c_int32),
        ("limit", C.c_int64),
        ("offset", C.c_int64),
    ]


class TpchDims(C.Structure):
    _fields_ = [("n_customer", C.c_int64), ("n_orders", C.c_int64), ("flags_mode", C.c_int32), ("reserved", C.c_int32)]


P = C.POINTER
_SIGNATURES = {
    "abi_version": (C.c_int, []),
    "last_error": (C.c_char_p, []),
    "kernel_launches": (C.c_int64, []),
    "create_hashes": (C.c_int, [P(ArrowArray), P(ArrowSchema), P(C.c_uint64)]),
    "eval_expr": (C.c_int, [P(Expr), P(Options), P(ArrowArray), P(ArrowSchema), P(ArrowArray), P(ArrowSchema)]),
    "filter_create": (C.c_int, [P(Expr), P(Options), P(C.c_void_p)]),
    "filter_execute": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema), P(ArrowArray), P(ArrowSchema)]),
    "filter_destroy": (None, [C.c_void_p]),
    "simple_agg_create": (C.c_int, [P(AggDesc), C.c_int32, P(Options), P(C.c_void_p)]),
    "simple_agg_push": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema)]),
    "simple_agg_finish": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema)]),
    "simple_agg_destroy": (None, [C.c_void_p]),
    "hash_agg_create": (C.c_int, [P(AggDesc), C.c_int32, P(Expr), P(C.c_char_p), C.c_int32, P(Options), P(C.c_void_p)]),
    "hash_agg_push": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema)]),
    "hash_agg_finish": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema)]),
    "hash_agg_destroy": (None, [C.c_void_p]),
    "hash_join_create": (C.c_int, [C.c_int32, P(Expr), P(Expr), C.c_int32, P(Expr), P(ArrowSchema), P(Options), P(C.c_void_p)]),
    "hash_join_build_push": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema)]),
    "hash_join_probe": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema), P(ArrowArray), P(ArrowSchema), P(C.c_int32)]),
    "hash_join_finish": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema), P(C.c_int32)]),
    "hash_join_destroy": (None, [C.c_void_p]),
    "cross_join_create": (C.c_int, [P(ArrowSchema), P(Options), P(C.c_void_p)]),
    "cross_join_build_push": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema)]),
    "cross_join_probe": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema)]),
    "cross_join_next": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema), P(C.c_int32)]),
    "cross_join_destroy": (None, [C.c_void_p]),
    "project_create": (C.c_int, [P(Expr), P(C.c_char_p), C.c_int32, P(Options), P(C.c_void_p)]),
    "project_execute": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema), P(ArrowArray), P(ArrowSchema)]),
    "project_destroy": (None, [C.c_void_p]),
    "order_create": (C.c_int, [P(Expr), P(C.c_int32), C.c_int32, P(Options), P(C.c_void_p)]),
    "order_push": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema)]),
    "order_finish": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema)]),
    "order_destroy": (None, [C.c_void_p]),
    "limit_create": (C.c_int, [C.c_int64, C.c_int64, P(Options), P(C.c_void_p)]),
    "limit_push": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema), P(ArrowArray), P(ArrowSchema), P(C.c_int32), P(C.c_int32)]),
    "limit_destroy": (None, [C.c_void_p]),
    "plan_create": (C.c_int, [P(PlanNode), C.c_int32, C.c_int32, P(Options), P(C.c_void_p)]),
    "plan_push_table": (C.c_int, [C.c_void_p, C.c_int32, P(ArrowArray), P(ArrowSchema)]),
    "plan_push_table_device": (C.c_int, [C.c_void_p, C.c_int32, P(ArrowDeviceArray), P(ArrowSchema)]),
    "table_create": (C.c_int, [P(Options), P(C.c_void_p)]),
    "table_append": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema)]),
    "table_num_rows": (C.c_int64, [C.c_void_p]),
    "table_num_batches": (C.c_int32, [C.c_void_p]),
    "table_read": (C.c_int, [C.c_void_p, C.c_int32, P(C.c_int32), C.c_int32, P(ArrowArray), P(ArrowSchema), P(C.c_int32)]),
    "table_destroy": (None, [C.c_void_p]),
    "table_read_csv": (C.c_int, [C.c_char_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, P(C.c_int32), C.c_int32, P(Options), P(C.c_void_p)]),
    "plan_push_table_resident": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "plan_execute": (C.c_int, [C.c_void_p]),
    "plan_next": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema), P(C.c_int32)]),
    "plan_reset": (C.c_int, [C.c_void_p]),
    "plan_clear_table": (C.c_int, [C.c_void_p, C.c_int32]),
    "plan_describe": (C.c_char_p, [C.c_void_p]),
    "plan_destroy": (None, [C.c_void_p]),
    "plan_execute_partial": (C.c_int, [C.c_void_p, C.c_int64]),
    "plan_export_partials": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema)]),
    "plan_clear_partials": (C.c_int, [C.c_void_p]),
    "plan_merge_partials": (C.c_int, [C.c_void_p, P(ArrowArray), P(ArrowSchema)]),
    "plan_finish_partial": (C.c_int, [C.c_void_p]),
    "plan_partials_row_words": (C.c_int, [C.c_void_p, P(C.c_int32)]),
    "plan_partials_tables": (C.c_int, [C.c_void_p, P(C.c_int32)]),
    "plan_select_partials_table": (C.c_int, [C.c_void_p, C.c_int32]),
    "plan_export_partials_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "plan_merge_partials_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64]),
    "plan_scan_kernel_ms": (C.c_double, [C.c_void_p, P(C.c_int64)]),
    "plan_push_table_batched": (C.c_int, [C.c_void_p, C.c_int32, P(ArrowArray), P(ArrowSchema), C.c_int64]),
    "plan_result_shape": (C.c_int, [C.c_void_p, P(C.c_int64), P(C.c_int32), P(C.c_int32)]),
    "plan_next_to_device": (C.c_int, [C.c_void_p, P(C.c_void_p), C.c_int32]),
    "plan_export_partials_partitioned": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, P(C.c_int64)]),
    "kernel_events_collect": (C.c_int, [P(C.c_void_p)]),
    "tpch_num_columns": (C.c_int32, [C.c_int32]),
    "tpch_num_rows": (C.c_int64, [P(TpchDims), C.c_int32]),
    "tpch_generate": (C.c_int, [P(TpchDims), C.c_int32, C.c_int64, C.c_int64, P(C.c_void_p), C.c_void_p]),
    "debug_compile_agg": (C.c_int, [P(AggDesc), C.c_int32, P(Expr), C.c_int32, P(Expr), P(ArrowSchema), P(Options), C.c_int32, P(C.c_void_p)]),
    "debug_compile_joinagg": (C.c_int, [P(AggDesc), C.c_int32, P(Expr), C.c_int32, P(Expr), C.c_int32, P(Expr), P(Expr), P(ArrowSchema), P(ArrowSchema),
                              P(Options), C.c_int32, P(C.c_void_p)]),
    "debug_compile_joinprobe": (C.c_int, [P(Expr), C.c_int32, P(Expr), P(ArrowSchema), P(Options), C.c_int32, P(C.c_void_p)]),
    "debug_compile_joinchain": (C.c_int, [P(Expr), C.c_int32, P(Expr), P(Expr), P(ArrowSchema), P(ArrowSchema), P(Options), C.c_int32, P(C.c_void_p)]),
    "debug_compile_eval": (C.c_int, [P(Expr), C.c_int32, C.c_int32, P(ArrowSchema), C.c_int32, P(C.c_void_p)]),
    "free": (None, [C.c_void_p]),
}
ABI_SYMBOLS = tuple(_SIGNATURES)


class Library:
    """One loaded build of the ABI."""

    def __init__(self, path: str, prefix: str = "sqlrs_"):
        self.path = path
        self.prefix = prefix
        self.cdll = C.CDLL(path, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(self.cdll, prefix + name)
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)
        if self.abi_version() != ABI_VERSION:
            raise RuntimeError(f"{path}: ABI version {self.abi_version()} != {ABI_VERSION}")

    def check(self, status: int):
        if status != OK:
            msg = self.last_error()
            raise ExecutorError(status, msg.decode() if msg else "")

    def options(self, count_mode=COUNT_REFERENCE_OVERWRITE, match_mode=MATCH_HASH_ONLY, device_id=-1, flags=0, stream=None):
        return Options(count_mode, match_mode, device_id, flags, stream)


_PRODUCT = None


def library_path() -> str:
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "csrc", "libsqlrs_b200.so")


def load() -> Library:
    """The CUDA library.  Fails loudly when it has not been built (no fallback of any kind)."""
    global _PRODUCT
    if _PRODUCT is None:
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing — run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a)")
        _PRODUCT = Library(path, "sqlrs_")
    return _PRODUCT


# ---- RecordBatch <-> C data ----------------------------------------------------------------------
def export_batch(batch: pa.RecordBatch):
    """-> (ArrowArray, ArrowSchema) owned by ctypes structs; the callee moves the array."""
    arr, sch = ArrowArray(), ArrowSchema()
    batch._export_to_c(C.addressof(arr), C.addressof(sch))
    return arr, sch


def export_schema(schema: pa.Schema) -> ArrowSchema:
    sch = ArrowSchema()
    schema._export_to_c(C.addressof(sch))
    return sch


def import_batch(arr: ArrowArray, sch: ArrowSchema) -> pa.RecordBatch:
    return pa.RecordBatch._import_from_c(C.addressof(arr), C.addressof(sch))


_RELEASE_SCHEMA = C.CFUNCTYPE(None, P(ArrowSchema))


def release_schema(sch: ArrowSchema):
    if sch.release:
        _RELEASE_SCHEMA(sch.release)(C.byref(sch))


def f64_bits(x: float) -> int:
    return struct.unpack("<q", struct.pack("<d", float(x)))[0]
