"""No-GPU checks of the drop-in boundary: the CUDA library loads, exports every symbol that
include/sqlrs_b200.h declares, and its run-time kernel specialisation compiles for sm_100a (NVRTC
needs no device).  No compute call is made here."""
import ctypes as C
import os
import re

import pyarrow as pa
import pytest

from sqlrs_b200.host import ffi, tpch
from sqlrs_b200.host.expr import AggArray, BinaryOp, Constant, ExprArray, InputRef, bind_binary_op

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "sqlrs_b200.h")).read()
    return sorted(set(re.findall(r"SQLRS_API\((\w+)\)\(", text)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(ffi.ABI_SYMBOLS)


def test_cuda_library_exports_every_declared_symbol(cuda_lib):
    for name in header_symbols():
        assert hasattr(cuda_lib.cdll, "sqlrs_" + name), name
    assert cuda_lib.abi_version() == ffi.ABI_VERSION == 2
    assert cuda_lib.kernel_launches() == 0 or cuda_lib.kernel_launches() > 0


def test_oracle_exports_the_same_abi(oracle):
    for name in header_symbols():
        assert hasattr(oracle.cdll, "sqlrs_oracle_" + name), name


def test_product_does_not_link_or_load_the_oracle(cuda_lib):
    import subprocess

    out = subprocess.run(["ldd", cuda_lib.path], capture_output=True, text=True).stdout
    assert "liboracle" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", cuda_lib.path], capture_output=True, text=True).stdout
    assert "sqlrs_oracle_" not in syms
    for root, _, files in os.walk(os.path.join(ROOT, "sqlrs_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".hpp", ".cuh")) and "embedded_sources" not in f:
                text = open(os.path.join(root, f)).read()
                assert "liboracle" not in text and "oracle/" not in text.replace("the oracle/", ""), os.path.join(root, f)


def _compile_agg(lib, plan, schema, count_mode, match_mode, compile_=1):
    aggs = AggArray(plan.agg_funcs, schema)
    groups = ExprArray(plan.group_by)
    pred = plan.child.expr.flatten()
    sch = ffi.export_schema(schema)
    src = C.c_void_p()
    opt = lib.options(count_mode=count_mode, match_mode=match_mode)
    try:
        lib.check(lib.debug_compile_agg(aggs.ptr, aggs.n, groups.ptr, groups.n, C.byref(pred.c), C.byref(sch), C.byref(opt), compile_, C.byref(src)))
        text = C.string_at(src.value).decode()
    finally:
        ffi.release_schema(sch)
        if src.value:
            lib.free(src)
    return text


@pytest.mark.parametrize("modes", [(ffi.COUNT_SQL_ACCUMULATE, ffi.MATCH_HASH_AND_KEY), (ffi.COUNT_REFERENCE_OVERWRITE, ffi.MATCH_HASH_ONLY)])
def test_q1_kernels_specialise_and_compile_for_sm100a(cuda_lib, modes):
    plan, schemas = tpch.q1_plan()
    text = _compile_agg(cuda_lib, plan, schemas[0], *modes)
    for needle in ("sq_agg_small", "sq_agg_global", "sq_agg_merge", "sq_agg_fixkeys", "__dmul_rn", "sq_row("):
        assert needle in text
    # every referenced column is loaded exactly once per row (common subexpressions are shared)
    assert text.count("SQ_LD_F64(1, r)") == 1 and text.count("SQ_LD_F64(2, r)") == 1
    if modes[1] == ffi.MATCH_HASH_ONLY:
        assert "sq_hash_one" in text.split("// ---- generated")[1].split("// ---- skeleton")[0]


def test_nullable_schema_compiles(cuda_lib):
    plan, _ = tpch.q1_plan()
    schema = pa.schema([pa.field(f.name, f.type, nullable=True) for f in tpch.schema_of(tpch.LINEITEM, tpch.Q1_COLUMNS)])
    text = _compile_agg(cuda_lib, plan, schema, ffi.COUNT_SQL_ACCUMULATE, ffi.MATCH_HASH_AND_KEY)
    assert "SQ_VALID(" in text


def test_eval_program_compiles_and_type_errors_surface_without_a_gpu(cuda_lib):
    schema = pa.schema([pa.field("a", pa.int64()), pa.field("b", pa.float64()), pa.field("c", pa.bool_())])
    exprs = ExprArray([bind_binary_op(InputRef(0, ffi.DT_INT64), "+", Constant(1)),
                       BinaryOp("AND", InputRef(2, ffi.DT_BOOL), BinaryOp(">", InputRef(1, ffi.DT_FLOAT64), Constant(0.5), ffi.DT_BOOL), ffi.DT_BOOL)])
    sch = ffi.export_schema(schema)
    src = C.c_void_p()
    try:
        cuda_lib.check(cuda_lib.debug_compile_eval(exprs.ptr, exprs.n, 0, C.byref(sch), 1, C.byref(src)))
        assert "sq_eval_kernel" in C.string_at(src.value).decode()
        cuda_lib.free(src)
        bad = ExprArray([BinaryOp("+", InputRef(0, ffi.DT_INT64), InputRef(1, ffi.DT_FLOAT64), ffi.DT_INT64)])
        with pytest.raises(ffi.ExecutorError) as err:
            cuda_lib.check(cuda_lib.debug_compile_eval(bad.ptr, bad.n, 0, C.byref(sch), 0, None))
        assert err.value.code == ffi.ERR_INTERNAL
        utf8 = ffi.export_schema(pa.schema([pa.field("s", pa.utf8())]))
        # Utf8 columns are string-pool ids on the device: equality compares the ids, an ORDERING comparison of strings their
        # byte-wise ranks (csrc/jit/strrank.cuh is compiled in only then)
        one = ExprArray([BinaryOp("=", InputRef(0, ffi.DT_UTF8), Constant("CO"), ffi.DT_BOOL)])
        cuda_lib.check(cuda_lib.debug_compile_eval(one.ptr, one.n, 0, C.byref(utf8), 1, C.byref(src)))
        assert "sq_rank_table" not in C.string_at(src.value).decode()
        cuda_lib.free(src)
        less = ExprArray([BinaryOp("<", InputRef(0, ffi.DT_UTF8), Constant("CO"), ffi.DT_BOOL)])
        cuda_lib.check(cuda_lib.debug_compile_eval(less.ptr, less.n, 0, C.byref(utf8), 1, C.byref(src)))
        text = C.string_at(src.value).decode()
        cuda_lib.free(src)
        assert "sq_rank_table" in text and "sq_str_rank(" in text
        ffi.release_schema(utf8)
    finally:
        ffi.release_schema(sch)


def test_generator_spec_row_counts(oracle):
    d = tpch.dims(0.01)
    assert tpch.num_rows(oracle, d, tpch.CUSTOMER) == 1500 and tpch.num_rows(oracle, d, tpch.ORDERS) == 15000
    n = tpch.num_rows(oracle, d, tpch.LINEITEM)
    t = tpch.host_table(oracle, d, tpch.LINEITEM)
    assert t.num_rows == n and 3.9 < n / 15000 < 4.1
    ok = t.column("l_orderkey").to_numpy()
    assert (ok[1:] >= ok[:-1]).all() and ok[0] == 1 and ok[-1] == 15000   # clustered on orderkey like dbgen


def test_q3_fused_probe_aggregate_kernel_compiles(cuda_lib):
    """the fused scan -> filter -> probe -> group-by kernel of Q3's last stage specialises and compiles for sm_100a"""
    from sqlrs_b200.host.expr import EMPTY_EXPR

    (_, _), (s2, s2s) = tpch.q3_stage_plans()
    join = s2.child
    aggs, groups = AggArray(s2.agg_funcs, join.join_output_schema), ExprArray(s2.group_by)
    rk = ExprArray([r for _, r in join.join_condition.on])
    pp = join.right.expr.flatten()
    bs, ps = ffi.export_schema(s2s[0]), ffi.export_schema(s2s[1])
    src = C.c_void_p()
    opt = cuda_lib.options(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    try:
        cuda_lib.check(cuda_lib.debug_compile_joinagg(aggs.ptr, aggs.n, groups.ptr, groups.n, rk.ptr, rk.n, C.byref(pp.c), C.byref(EMPTY_EXPR.c),
                                                      C.byref(bs), C.byref(ps), C.byref(opt), 1, C.byref(src)))
        text = C.string_at(src.value).decode()
    finally:
        ffi.release_schema(bs)
        ffi.release_schema(ps)
        if src.value:
            cuda_lib.free(src)
    assert "sq_joinagg_kernel" in text and "sq_probe_row" in text and "SQ_LDB_I64(4, b)" in text


def test_q3_fused_probe_kernel_compiles(cuda_lib):
    """the fused scan -> filter -> key hash -> Bloom -> probe kernel of Q3's first join (orders probing customer)
    specialises and compiles for sm_100a, in both identity modes"""
    plan, schemas = tpch.q3_plan()
    j1 = plan.child.left
    rk = ExprArray([r for _, r in j1.join_condition.on])
    pp = j1.right.expr.flatten()
    ps = ffi.export_schema(schemas[1])
    try:
        for mm in (ffi.MATCH_HASH_AND_KEY, ffi.MATCH_HASH_ONLY):
            src = C.c_void_p()
            opt = cuda_lib.options(match_mode=mm)
            cuda_lib.check(cuda_lib.debug_compile_joinprobe(rk.ptr, rk.n, C.byref(pp.c), C.byref(ps), C.byref(opt), 1, C.byref(src)))
            text = C.string_at(src.value).decode()
            cuda_lib.free(src)
            assert "sq_joinprobe_kernel" in text and "sq_probe_row" in text and f"#define SQ_JMATCH {mm}" in text
    finally:
        ffi.release_schema(ps)
