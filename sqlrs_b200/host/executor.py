"""Host-side mirror of the reference's operator interface (src/executor/mod.rs:34-64).

The reference's boundary is "a struct of plan parameters + child stream(s) with
`execute() -> BoxStream<Result<RecordBatch, ExecutorError>>`".  The classes below have the
same names and fields (`FilterExecutor{expr, child}`, `SimpleAggExecutor{agg_funcs, child}`,
`HashAggExecutor{agg_funcs, group_by, child}`, `HashJoinExecutor{left_child, right_child,
join_type, join_condition, join_output_schema}`, and the operators that follow them in a plan:
`ProjectExecutor{exprs, child}`, `OrderExecutor{order_by, child}`, `LimitExecutor{limit, offset,
child}`); `child` is any iterable of
`pyarrow.RecordBatch`, `execute()` is a generator of `RecordBatch`, errors surface as
`ExecutorError`.  Every batch crosses the C ABI of include/sqlrs_b200.h through the Arrow C
Data Interface — this file contains no compute.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import pyarrow as pa

from . import ffi
from .expr import AggArray, BoundExpr, ExprArray, FlatExpr, NameArray, EMPTY_EXPR, project_field_names

BoxedExecutor = Iterable[pa.RecordBatch]


def try_collect(executor: BoxedExecutor) -> List[pa.RecordBatch]:
    """src/executor/mod.rs:58-64"""
    return list(executor)


def _lib(lib):
    return lib if lib is not None else ffi.load()


def _import(arr, sch):
    return ffi.import_batch(arr, sch)


class FilterExecutor:
    """src/executor/filter.rs:7-26"""

    def __init__(self, expr: BoundExpr, child: BoxedExecutor, lib: Optional[ffi.Library] = None, options=None):
        self.expr, self.child, self.lib = expr, child, _lib(lib)
        self.options = options if options is not None else self.lib.options()

    def execute(self) -> Iterator[pa.RecordBatch]:
        lib = self.lib
        flat = self.expr.flatten()
        h = C.c_void_p()
        lib.check(lib.filter_create(C.byref(flat.c), C.byref(self.options), C.byref(h)))
        try:
            for batch in self.child:
                arr, sch = ffi.export_batch(batch)
                out, out_sch = ffi.ArrowArray(), ffi.ArrowSchema()
                try:
                    lib.check(lib.filter_execute(h, C.byref(arr), C.byref(sch), C.byref(out), C.byref(out_sch)))
                finally:
                    ffi.release_schema(sch)
                yield _import(out, out_sch)
        finally:
            lib.filter_destroy(h)


class SimpleAggExecutor:
    """src/executor/aggregate/simple_agg.rs:10-65"""

    def __init__(self, agg_funcs: Sequence[BoundExpr], child: BoxedExecutor, lib=None, options=None):
        self.agg_funcs, self.child, self.lib = list(agg_funcs), child, _lib(lib)
        self.options = options if options is not None else self.lib.options()

    def execute(self) -> Iterator[pa.RecordBatch]:
        lib = self.lib
        h = None
        keep = None
        try:
            for batch in self.child:
                if h is None:
                    keep = AggArray(self.agg_funcs, batch.schema)
                    h = C.c_void_p()
                    lib.check(lib.simple_agg_create(keep.ptr, keep.n, C.byref(self.options), C.byref(h)))
                arr, sch = ffi.export_batch(batch)
                try:
                    lib.check(lib.simple_agg_push(h, C.byref(arr), C.byref(sch)))
                finally:
                    ffi.release_schema(sch)
            if h is None:  # simple_agg.rs:63 unwraps None when the child yields nothing
                raise ffi.ExecutorError(ffi.ERR_INTERNAL, "called `Option::unwrap()` on a `None` value (no input batch)")
            out, out_sch = ffi.ArrowArray(), ffi.ArrowSchema()
            lib.check(lib.simple_agg_finish(h, C.byref(out), C.byref(out_sch)))
            yield _import(out, out_sch)
        finally:
            if h is not None:
                lib.simple_agg_destroy(h)


class HashAggExecutor:
    """src/executor/aggregate/hash_agg.rs:15-150"""

    def __init__(self, agg_funcs: Sequence[BoundExpr], group_by: Sequence[BoundExpr], child: BoxedExecutor, lib=None, options=None):
        self.agg_funcs, self.group_by, self.child, self.lib = list(agg_funcs), list(group_by), child, _lib(lib)
        self.options = options if options is not None else self.lib.options()

    def execute(self) -> Iterator[pa.RecordBatch]:
        lib = self.lib
        h = None
        keep = None
        try:
            for batch in self.child:
                if h is None:
                    aggs = AggArray(self.agg_funcs, batch.schema)
                    groups = ExprArray(self.group_by)
                    names = NameArray([g.eval_field(batch.schema).name for g in self.group_by])
                    keep = (aggs, groups, names)
                    h = C.c_void_p()
                    lib.check(lib.hash_agg_create(aggs.ptr, aggs.n, groups.ptr, names.ptr, groups.n, C.byref(self.options), C.byref(h)))
                arr, sch = ffi.export_batch(batch)
                try:
                    lib.check(lib.hash_agg_push(h, C.byref(arr), C.byref(sch)))
                finally:
                    ffi.release_schema(sch)
            if h is None:  # hash_agg.rs:125
                raise ffi.ExecutorError(ffi.ERR_INTERNAL, "called `Option::unwrap()` on a `None` value (no input batch)")
            out, out_sch = ffi.ArrowArray(), ffi.ArrowSchema()
            lib.check(lib.hash_agg_finish(h, C.byref(out), C.byref(out_sch)))
            yield _import(out, out_sch)
        finally:
            if h is not None:
                lib.hash_agg_destroy(h)


@dataclass
class JoinCondition:
    """JoinCondition::On{on, filter} (src/binder/table/join.rs:26-48)"""
    on: List[Tuple[BoundExpr, BoundExpr]]
    filter: Optional[BoundExpr] = None


JOIN_TYPES = {"Inner": ffi.JOIN_INNER, "Left": ffi.JOIN_LEFT, "Right": ffi.JOIN_RIGHT, "Full": ffi.JOIN_FULL}


class HashJoinExecutor:
    """src/executor/join/hash_join.rs:16-23,147-323.  `join_output_schema` is the pyarrow schema of
    the joined row (all left fields then all right fields, catalog/mod.rs:131-138)."""

    def __init__(self, left_child, right_child, join_type: str, join_condition: JoinCondition, join_output_schema: pa.Schema,
                 lib=None, options=None):
        self.left_child, self.right_child = left_child, right_child
        self.join_type, self.join_condition, self.join_output_schema = join_type, join_condition, join_output_schema
        self.lib = _lib(lib)
        self.options = options if options is not None else self.lib.options()

    def execute(self) -> Iterator[pa.RecordBatch]:
        lib = self.lib
        if self.join_type not in JOIN_TYPES:
            raise ffi.ExecutorError(ffi.ERR_INTERNAL, "Cross join should not be in HashJoinExecutor")
        lk = ExprArray([l for l, _ in self.join_condition.on])
        rk = ExprArray([r for _, r in self.join_condition.on])
        flt = self.join_condition.filter.flatten() if self.join_condition.filter is not None else EMPTY_EXPR
        out_schema_c = ffi.export_schema(self.join_output_schema)
        h = C.c_void_p()
        try:
            lib.check(lib.hash_join_create(JOIN_TYPES[self.join_type], lk.ptr, rk.ptr, lk.n, C.byref(flt.c), C.byref(out_schema_c),
                                           C.byref(self.options), C.byref(h)))
        finally:
            ffi.release_schema(out_schema_c)
        try:
            for batch in self.left_child:
                arr, sch = ffi.export_batch(batch)
                try:
                    lib.check(lib.hash_join_build_push(h, C.byref(arr), C.byref(sch)))
                finally:
                    ffi.release_schema(sch)
            has = C.c_int32(0)
            for batch in self.right_child:
                arr, sch = ffi.export_batch(batch)
                out, out_sch = ffi.ArrowArray(), ffi.ArrowSchema()
                try:
                    lib.check(lib.hash_join_probe(h, C.byref(arr), C.byref(sch), C.byref(out), C.byref(out_sch), C.byref(has)))
                finally:
                    ffi.release_schema(sch)
                if has.value:
                    yield _import(out, out_sch)
            out, out_sch = ffi.ArrowArray(), ffi.ArrowSchema()
            lib.check(lib.hash_join_finish(h, C.byref(out), C.byref(out_sch), C.byref(has)))
            if has.value:
                yield _import(out, out_sch)
        finally:
            lib.hash_join_destroy(h)


class CrossJoinExecutor:
    """src/executor/join/cross_join.rs:8-57: one output batch per (right batch, left row)"""

    def __init__(self, left_child, right_child, join_output_schema: pa.Schema, lib=None, options=None):
        self.left_child, self.right_child, self.join_output_schema = left_child, right_child, join_output_schema
        self.lib = _lib(lib)
        self.options = options if options is not None else self.lib.options()

    def execute(self) -> Iterator[pa.RecordBatch]:
        lib = self.lib
        out_schema_c = ffi.export_schema(self.join_output_schema)
        h = C.c_void_p()
        try:
            lib.check(lib.cross_join_create(C.byref(out_schema_c), C.byref(self.options), C.byref(h)))
        finally:
            ffi.release_schema(out_schema_c)
        try:
            for batch in self.left_child:
                arr, sch = ffi.export_batch(batch)
                try:
                    lib.check(lib.cross_join_build_push(h, C.byref(arr), C.byref(sch)))
                finally:
                    ffi.release_schema(sch)
            has = C.c_int32(0)
            for batch in self.right_child:
                arr, sch = ffi.export_batch(batch)
                try:
                    lib.check(lib.cross_join_probe(h, C.byref(arr), C.byref(sch)))
                finally:
                    ffi.release_schema(sch)
                while True:
                    out, out_sch = ffi.ArrowArray(), ffi.ArrowSchema()
                    lib.check(lib.cross_join_next(h, C.byref(out), C.byref(out_sch), C.byref(has)))
                    if not has.value:
                        break
                    yield _import(out, out_sch)
        finally:
            lib.cross_join_destroy(h)


class ProjectExecutor:
    """src/executor/project.rs:6-29"""

    def __init__(self, exprs: Sequence[BoundExpr], child: BoxedExecutor, lib=None, options=None):
        self.exprs, self.child, self.lib = list(exprs), child, _lib(lib)
        self.options = options if options is not None else self.lib.options()

    def execute(self) -> Iterator[pa.RecordBatch]:
        lib = self.lib
        h = None
        keep = None
        try:
            for batch in self.child:
                if h is None:
                    exprs = ExprArray(self.exprs)
                    names = NameArray(project_field_names(self.exprs, batch.schema))
                    keep = (exprs, names)
                    h = C.c_void_p()
                    lib.check(lib.project_create(exprs.ptr, names.ptr, exprs.n, C.byref(self.options), C.byref(h)))
                arr, sch = ffi.export_batch(batch)
                out, out_sch = ffi.ArrowArray(), ffi.ArrowSchema()
                try:
                    lib.check(lib.project_execute(h, C.byref(arr), C.byref(sch), C.byref(out), C.byref(out_sch)))
                finally:
                    ffi.release_schema(sch)
                yield _import(out, out_sch)
        finally:
            if h is not None:
                lib.project_destroy(h)


@dataclass
class BoundOrderBy:
    """BoundOrderBy{expr, asc} (src/binder/statement/select.rs)"""
    expr: BoundExpr
    asc: bool = True


class OrderExecutor:
    """src/executor/order.rs:8-67"""

    def __init__(self, order_by: Sequence[BoundOrderBy], child: BoxedExecutor, lib=None, options=None):
        self.order_by, self.child, self.lib = list(order_by), child, _lib(lib)
        self.options = options if options is not None else self.lib.options()

    def execute(self) -> Iterator[pa.RecordBatch]:
        lib = self.lib
        exprs = ExprArray([o.expr for o in self.order_by])
        asc = (C.c_int32 * max(1, len(self.order_by)))(*[int(o.asc) for o in self.order_by])
        h = C.c_void_p()
        lib.check(lib.order_create(exprs.ptr, asc, exprs.n, C.byref(self.options), C.byref(h)))
        try:
            for batch in self.child:
                arr, sch = ffi.export_batch(batch)
                try:
                    lib.check(lib.order_push(h, C.byref(arr), C.byref(sch)))
                finally:
                    ffi.release_schema(sch)
            out, out_sch = ffi.ArrowArray(), ffi.ArrowSchema()
            lib.check(lib.order_finish(h, C.byref(out), C.byref(out_sch)))  # zero batches: order.rs:27 unwraps None
            yield _import(out, out_sch)
        finally:
            lib.order_destroy(h)


class LimitExecutor:
    """src/executor/limit.rs:6-80; `limit` / `offset` are the bound constants or None"""

    def __init__(self, limit: Optional[int], offset: Optional[int], child: BoxedExecutor, lib=None, options=None):
        self.limit, self.offset, self.child, self.lib = limit, offset, child, _lib(lib)
        self.options = options if options is not None else self.lib.options()

    def execute(self) -> Iterator[pa.RecordBatch]:
        lib = self.lib
        if self.limit is not None and self.limit == 0:  # limit.rs:31-33: returns before polling the child
            return
        h = C.c_void_p()
        lib.check(lib.limit_create(-1 if self.limit is None else self.limit, -1 if self.offset is None else self.offset,
                                   C.byref(self.options), C.byref(h)))
        try:
            has, done = C.c_int32(0), C.c_int32(0)
            for batch in self.child:
                arr, sch = ffi.export_batch(batch)
                out, out_sch = ffi.ArrowArray(), ffi.ArrowSchema()
                try:
                    lib.check(lib.limit_push(h, C.byref(arr), C.byref(sch), C.byref(out), C.byref(out_sch), C.byref(has), C.byref(done)))
                finally:
                    ffi.release_schema(sch)
                if has.value:
                    yield _import(out, out_sch)
                if done.value:
                    break
        finally:
            lib.limit_destroy(h)


def eval_column(expr: BoundExpr, batch: pa.RecordBatch, lib=None, options=None) -> pa.Array:
    """BoundExpr::eval_column (src/executor/evaluator.rs:13-28)"""
    lib = _lib(lib)
    options = options if options is not None else lib.options()
    flat = expr.flatten()
    arr, sch = ffi.export_batch(batch)
    out, out_sch = ffi.ArrowArray(), ffi.ArrowSchema()
    try:
        lib.check(lib.eval_expr(C.byref(flat.c), C.byref(options), C.byref(arr), C.byref(sch), C.byref(out), C.byref(out_sch)))
    finally:
        ffi.release_schema(sch)
    return _import(out, out_sch).column(0)


def create_hashes(arrays: Sequence[pa.Array], lib=None):
    """create_hashes (src/executor/aggregate/hash_utils.rs:161-220), seeds (0,0,0,0) -> list of u64"""
    lib = _lib(lib)
    batch = pa.RecordBatch.from_arrays(list(arrays), names=[f"c{i}" for i in range(len(arrays))])
    arr, sch = ffi.export_batch(batch)
    out = (C.c_uint64 * max(1, batch.num_rows))()
    try:
        lib.check(lib.create_hashes(C.byref(arr), C.byref(sch), out))
    finally:
        ffi.release_schema(sch)
    return list(out)[: batch.num_rows]
