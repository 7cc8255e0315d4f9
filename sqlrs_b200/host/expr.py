"""Host-side mirror of the reference's bound expressions (src/binder/expression/mod.rs:18-27).

Same names and argument meaning as the Rust enum variants — `InputRef`, `Constant`,
`BinaryOp`, `TypeCast`, `Alias`, `AggFunc` — plus `bind_binary_op`, which applies the binder's
implicit widening casts (src/binder/expression/binary_op.rs:27-73).  `flatten()` turns a tree
into the postfix bytecode of include/sqlrs_b200.h; `eval_field()` reproduces the output field
names of src/executor/evaluator.rs:30-64.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import pyarrow as pa

from . import ffi

_OPS = {
    "+": ffi.OP_ADD, "-": ffi.OP_SUB, "*": ffi.OP_MUL, "/": ffi.OP_DIV,
    ">": ffi.OP_GT, "<": ffi.OP_LT, ">=": ffi.OP_GE, "<=": ffi.OP_LE, "=": ffi.OP_EQ, "<>": ffi.OP_NE,
    "AND": ffi.OP_AND, "OR": ffi.OP_OR,
    # the v2 engine's checked arithmetic (src/function/scalar/arithmetic_function.rs): integer overflow is an error
    "+checked": ffi.OP_ADD_CHECKED, "-checked": ffi.OP_SUB_CHECKED, "*checked": ffi.OP_MUL_CHECKED, "/checked": ffi.OP_DIV_CHECKED,
}
_ARITH = {"+", "-", "*", "/", "+checked", "-checked", "*checked", "/checked"}


def fold_and(predicates):
    """The v2 Filter's predicate: its expressions folded into one AND conjunction, left to right
    (BoundConjunctionExpression::try_build_and_conjunction_expression, src/execution/physical_plan/physical_filter.rs:11-16)."""
    predicates = list(predicates)
    if not predicates:
        raise ValueError("fold_and: a Filter has at least one predicate")
    out = predicates[0]
    for p in predicates[1:]:
        out = BinaryOp("AND", out, p, ffi.DT_BOOL)
    return out


class BoundExpr:
    def return_type(self) -> Optional[int]:
        raise NotImplementedError

    def eval_field(self, schema: pa.Schema) -> pa.Field:
        raise NotImplementedError

    def _emit(self, out: list):
        raise NotImplementedError

    def flatten(self) -> "FlatExpr":
        nodes: list = []
        self._emit(nodes)
        return FlatExpr(nodes)


@dataclass
class InputRef(BoundExpr):
    """BoundInputRef{index, return_type}"""
    index: int
    return_dtype: int = ffi.DT_INT64

    def return_type(self):
        return self.return_dtype

    def eval_field(self, schema):
        return schema.field(self.index)

    def _emit(self, out):
        out.append((ffi.OP_INPUT_REF, self.return_dtype, self.index, 0, 0, None))


@dataclass
class Constant(BoundExpr):
    """BoundExpr::Constant(ScalarValue); python ints bind like SQL literals (types/mod.rs:150-160)."""
    value: object
    dtype: Optional[int] = None

    def __post_init__(self):
        if self.dtype is None:
            v = self.value
            if v is None:
                self.dtype = ffi.DT_NULL
            elif isinstance(v, bool):
                self.dtype = ffi.DT_BOOL
            elif isinstance(v, int):
                self.dtype = ffi.DT_INT32 if -(2**31) <= v < 2**31 else ffi.DT_INT64
            elif isinstance(v, float):
                self.dtype = ffi.DT_FLOAT64
            elif isinstance(v, str):
                self.dtype = ffi.DT_UTF8
            else:
                raise TypeError(f"unsupported constant {v!r}")

    def return_type(self):
        return self.dtype

    def _display(self):
        if self.dtype == ffi.DT_NULL:
            return "Null"
        if self.value is None:
            return "NULL"
        if self.dtype == ffi.DT_BOOL:
            return "true" if self.value else "false"
        if self.dtype == ffi.DT_FLOAT64:
            v = float(self.value)
            return str(int(v)) if v.is_integer() and abs(v) < 1e16 else repr(v)
        return str(self.value)

    def eval_field(self, schema):
        return pa.field(self._display(), ffi.pa_type_of(self.dtype), True)

    def _emit(self, out):
        is_null = int(self.value is None)
        bits, s = 0, None
        if not is_null:
            if self.dtype == ffi.DT_FLOAT64:
                bits = ffi.f64_bits(self.value)
            elif self.dtype == ffi.DT_UTF8:
                s = str(self.value).encode()
            else:
                bits = int(self.value)
        out.append((ffi.OP_CONSTANT, self.dtype, 0, is_null, bits, s))


@dataclass
class TypeCast(BoundExpr):
    """BoundTypeCast{expr, cast_type}"""
    expr: BoundExpr
    cast_type: int

    def return_type(self):
        return self.cast_type

    def eval_field(self, schema):
        inner = self.expr.eval_field(schema)
        return pa.field(f"{ffi.DT_NAME[self.cast_type]}({inner.name})", ffi.pa_type_of(self.cast_type), True)

    def _emit(self, out):
        self.expr._emit(out)
        out.append((ffi.OP_CAST, self.cast_type, 0, 0, 0, None))


@dataclass
class BinaryOp(BoundExpr):
    """BoundBinaryOp{op, left, right, return_type}; op is the sqlparser spelling ('+', '<=', 'AND' ...)"""
    op: str
    left: BoundExpr
    right: BoundExpr
    return_dtype: Optional[int] = None

    def __post_init__(self):
        if self.op not in _OPS:
            raise ffi.ExecutorError(ffi.ERR_UNSUPPORTED, f"not supported binary operator: {self.op}")
        if self.return_dtype is None:
            self.return_dtype = self.left.return_type() if self.op in _ARITH else ffi.DT_BOOL

    def return_type(self):
        return self.return_dtype

    def eval_field(self, schema):
        l, r = self.left.eval_field(schema), self.right.eval_field(schema)
        return pa.field(f"{l.name}{self.op.replace('checked', '')}{r.name}", ffi.pa_type_of(self.return_dtype), True)

    def _emit(self, out):
        self.left._emit(out)
        self.right._emit(out)
        out.append((_OPS[self.op], self.return_dtype, 0, 0, 0, None))


@dataclass
class Alias(BoundExpr):
    """BoundAlias{expr, column_id}: evaluated as the inner expression (evaluator.rs:25)"""
    expr: BoundExpr
    column_id: str

    def return_type(self):
        return self.expr.return_type()

    def eval_field(self, schema):
        return pa.field(self.column_id, ffi.pa_type_of(self.expr.return_type()), True)

    def _emit(self, out):
        self.expr._emit(out)


_AGG = {"Count": ffi.AGG_COUNT, "Sum": ffi.AGG_SUM, "Min": ffi.AGG_MIN, "Max": ffi.AGG_MAX}


@dataclass
class AggFunc(BoundExpr):
    """BoundAggFunc{func, exprs, return_type, distinct}; return types as bind_agg_func
    (src/binder/expression/agg_func.rs:54-80): Count -> Int64, Sum/Min/Max -> type of exprs[0]."""
    func: str
    exprs: Sequence[BoundExpr]
    return_dtype: Optional[int] = None
    distinct: bool = False

    def __post_init__(self):
        if self.func not in _AGG:
            raise ffi.ExecutorError(ffi.ERR_UNSUPPORTED, f"not implmented agg func {self.func}")
        if self.return_dtype is None:
            self.return_dtype = ffi.DT_INT64 if self.func == "Count" else self.exprs[0].return_type()

    def return_type(self):
        return self.return_dtype

    def eval_field(self, schema):
        inner = self.exprs[0].eval_field(schema)
        return pa.field(f"{self.func}({inner.name})", ffi.pa_type_of(self.return_dtype), True)


def bind_binary_op(left: BoundExpr, op: str, right: BoundExpr) -> BinaryOp:
    """Binder::bind_binary_op (binary_op.rs:18-81): widen Int32 -> Int64 -> Float64 with TypeCast."""
    lt, rt = left.return_type(), right.return_type()
    ret = lt
    if lt != rt:
        I32, I64, F64 = ffi.DT_INT32, ffi.DT_INT64, ffi.DT_FLOAT64
        if (lt, rt) in ((I64, I32), (F64, I32), (F64, I64)):
            right = TypeCast(right, lt)
        elif (lt, rt) in ((I32, I64), (I32, F64), (I64, F64)):
            left = TypeCast(left, rt)
            ret = rt
        else:
            raise ffi.ExecutorError(ffi.ERR_UNSUPPORTED, "not implmented type conversion")
    return BinaryOp(op, left, right, ret if op in _ARITH else ffi.DT_BOOL)


# ---- bytecode holders (keep ctypes memory alive) ---------------------------------------------
class FlatExpr:
    def __init__(self, nodes: list):
        self.n = len(nodes)
        self._arr = (ffi.ExprNode * max(1, self.n))()
        for k, (op, dt, idx, is_null, bits, s) in enumerate(nodes):
            self._arr[k] = ffi.ExprNode(op, dt, idx, is_null, bits, s)
        self.c = ffi.Expr(C.cast(self._arr, C.POINTER(ffi.ExprNode)), self.n)


EMPTY_EXPR = FlatExpr([])


def project_field_names(exprs: Sequence[BoundExpr], schema: pa.Schema) -> List[Optional[str]]:
    """ProjectExecutor's output field names at the ABI: None = a bare InputRef keeps the input field
    (evaluator.rs:31); everything else (Alias included) is a new nullable field named by eval_field."""
    return [None if isinstance(e, InputRef) else e.eval_field(schema).name for e in exprs]


class ExprArray:
    def __init__(self, exprs: Sequence[BoundExpr]):
        self.flat = [e.flatten() for e in exprs]
        self._arr = (ffi.Expr * max(1, len(self.flat)))()
        for k, f in enumerate(self.flat):
            self._arr[k] = f.c
        self.ptr = C.cast(self._arr, C.POINTER(ffi.Expr))
        self.n = len(self.flat)


class AggArray:
    def __init__(self, agg_funcs: Sequence[BoundExpr], schema: Optional[pa.Schema], names: Optional[List[str]] = None):
        self.aggs = []
        for e in agg_funcs:
            while isinstance(e, Alias):
                e = e.expr
            if not isinstance(e, AggFunc):
                raise ffi.ExecutorError(ffi.ERR_INTERNAL, "create_accumulator called with non-aggregate expression")
            self.aggs.append(e)
        self.flat = [a.exprs[0].flatten() for a in self.aggs]
        if names is None:
            names = [e.eval_field(schema).name for e in agg_funcs]
        self._names = [n.encode() for n in names]
        self._arr = (ffi.AggDesc * max(1, len(self.aggs)))()
        for k, a in enumerate(self.aggs):
            self._arr[k] = ffi.AggDesc(_AGG[a.func], int(a.distinct), a.return_dtype, 0, self.flat[k].c, self._names[k])
        self.ptr = C.cast(self._arr, C.POINTER(ffi.AggDesc))
        self.n = len(self.aggs)


class NameArray:
    def __init__(self, names: Sequence[Optional[str]]):
        self._b = [None if n is None else n.encode() for n in names]
        self._arr = (C.c_char_p * max(1, len(self._b)))(*self._b)
        self.ptr = C.cast(self._arr, C.POINTER(C.c_char_p))
