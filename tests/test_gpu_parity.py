"""Parity of the CUDA path with the CPU oracle, through the C ABI, on seeded random inputs.
Integer / Boolean / index results bit-exact (row order included); Float64 sums within 1e-9 relative
(north-star tolerance: 1e-6).  Covers the edge cases the reference tests and the ones it leaves
unpinned: NULLs, empty and sliced batches, several batches per side, the quirks
K1/K2/K3 in both behaviour modes, hash-table growth, heavy key duplication."""
import ctypes as C

import numpy as np
import pyarrow as pa
import pytest

from sqlrs_b200.host import executor as ex
from sqlrs_b200.host import ffi, tpch
from sqlrs_b200.host.expr import AggFunc, BinaryOp, Constant, InputRef, TypeCast, bind_binary_op
from sqlrs_b200.host.plan import ExecutorBuilder
from util import assert_batches_match, rows_of

pytestmark = pytest.mark.gpu
I32, I64, F64, BOOL = ffi.DT_INT32, ffi.DT_INT64, ffi.DT_FLOAT64, ffi.DT_BOOL
PA = {I32: pa.int32(), I64: pa.int64(), F64: pa.float64(), BOOL: pa.bool_()}
FTOL = 1e-9


def random_batch(rng, n, dtypes, null_frac=0.15, small_ints=False, names=None):
    arrays = []
    for dt in dtypes:
        if dt == F64:
            v = np.round(rng.normal(0, 100, n), 2)
        elif dt == BOOL:
            v = rng.integers(0, 2, n).astype(bool)
        elif dt == I32:
            v = rng.integers(-6, 7, n).astype(np.int32) if small_ints else rng.integers(-2**31, 2**31 - 1, n).astype(np.int32)
        else:
            v = rng.integers(-6, 7, n).astype(np.int64) if small_ints else rng.integers(-2**62, 2**62, n).astype(np.int64)
        mask = rng.random(n) < null_frac if null_frac > 0 else None
        arrays.append(pa.array(v, type=PA[dt], mask=mask))
    names = names or [f"c{i}" for i in range(len(dtypes))]
    return pa.RecordBatch.from_arrays(arrays, names=names)


def both(fn, cuda_lib, oracle):
    return fn(cuda_lib), fn(oracle)


# ------------------------------------------------------------------ expressions
def random_expr(rng, dtypes, want, depth):
    """A random well-typed BoundExpr of result type `want` over columns of `dtypes`."""
    cols = [i for i, d in enumerate(dtypes) if d == want]
    if depth == 0 or rng.random() < 0.25:
        if cols and rng.random() < 0.8:
            return InputRef(int(rng.choice(cols)), want)
        if want == BOOL:
            return Constant(bool(rng.integers(0, 2)))
        if want == F64:
            return Constant(float(np.round(rng.normal(0, 10), 1)))
        v = int(rng.integers(-5, 6))
        if rng.random() < 0.1:
            return Constant(None, want)
        return Constant(v, want)
    if want == BOOL:
        kind = rng.choice(["cmp", "logic", "cast"])
        if kind == "cmp":
            t = int(rng.choice([I32, I64, F64, BOOL]))
            op = str(rng.choice([">", "<", ">=", "<=", "=", "<>"]))
            return BinaryOp(op, random_expr(rng, dtypes, t, depth - 1), random_expr(rng, dtypes, t, depth - 1), BOOL)
        if kind == "logic":
            return BinaryOp(str(rng.choice(["AND", "OR"])), random_expr(rng, dtypes, BOOL, depth - 1), random_expr(rng, dtypes, BOOL, depth - 1), BOOL)
        return TypeCast(random_expr(rng, dtypes, int(rng.choice([I32, I64, F64])), depth - 1), BOOL)
    if rng.random() < 0.25:
        src = int(rng.choice([t for t in (I32, I64, F64, BOOL) if t != want]))
        return TypeCast(random_expr(rng, dtypes, src, depth - 1), want)
    op = str(rng.choice(["+", "-", "*"]))
    return BinaryOp(op, random_expr(rng, dtypes, want, depth - 1), random_expr(rng, dtypes, want, depth - 1), want)


@pytest.mark.parametrize("seed", range(12))
def test_random_expressions(cuda_lib, oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    dtypes = [I32, I64, F64, BOOL, I64, F64, I32, BOOL]
    n = int(rng.choice([0, 1, 31, 32, 33, 1000, 4097]))
    b = random_batch(rng, n, dtypes, null_frac=float(rng.choice([0.0, 0.2])))
    for _ in range(6):
        want = int(rng.choice([I32, I64, F64, BOOL]))
        e = random_expr(rng, dtypes, want, 3)
        got, exp = both(lambda l: ex.eval_column(e, b, lib=l), cuda_lib, oracle)
        assert got.type == exp.type, (e, got.type, exp.type)
        assert got.to_pylist() == exp.to_pylist() or all(
            (g == w) or (g is not None and w is not None and g != g and w != w) for g, w in zip(got.to_pylist(), exp.to_pylist())), e


def test_integer_division_and_divide_by_zero(cuda_lib, oracle):
    b = pa.RecordBatch.from_arrays([pa.array([7, -7, -2**63, 5, None], pa.int64()), pa.array([2, 2, -1, None, 0], pa.int64())], names=["a", "b"])
    e = BinaryOp("/", InputRef(0, I64), InputRef(1, I64), I64)
    got, exp = both(lambda l: ex.eval_column(e, b, lib=l).to_pylist(), cuda_lib, oracle)
    assert got == exp == [3, -3, -2**63, None, None]
    bad = pa.RecordBatch.from_arrays([pa.array([1, 2], pa.int64()), pa.array([1, 0], pa.int64())], names=["a", "b"])
    for lib in (cuda_lib, oracle):
        with pytest.raises(ffi.ExecutorError) as err:
            ex.eval_column(e, bad, lib=lib)
        assert err.value.code == ffi.ERR_ARROW and "Divide by zero" in err.value.message


def test_type_errors_match_the_oracle(cuda_lib, oracle):
    b = random_batch(np.random.default_rng(5), 10, [I32, I64, F64, BOOL])
    cases = [BinaryOp("+", InputRef(0, I32), InputRef(1, I64), I64), BinaryOp(">", InputRef(0, I32), InputRef(2, F64), BOOL),
             BinaryOp("AND", InputRef(0, I32), InputRef(3, BOOL), BOOL), BinaryOp("+", InputRef(3, BOOL), InputRef(3, BOOL), BOOL),
             InputRef(9, I64)]
    for e in cases:
        codes = []
        for lib in (cuda_lib, oracle):
            with pytest.raises(ffi.ExecutorError) as err:
                ex.eval_column(e, b, lib=lib)
            codes.append(err.value.code)
        assert codes[0] == codes[1], (e, codes)


# ------------------------------------------------------------------ filter
@pytest.mark.parametrize("seed", range(6))
def test_filter_random_and_sliced(cuda_lib, oracle, seed):
    rng = np.random.default_rng(2000 + seed)
    dtypes = [I64, F64, I32, BOOL, I64]
    n = int(rng.choice([1, 63, 64, 65, 5000, 70001]))
    full = random_batch(rng, n + 37, dtypes, null_frac=0.2, small_ints=True)
    b = full.slice(int(rng.integers(0, 37)), n)  # non-zero offsets, unaligned bitmaps (Limit slices batches: limit.rs:69)
    pred = BinaryOp("OR", bind_binary_op(InputRef(0, I64), ">", Constant(0)),
                    BinaryOp("AND", InputRef(3, BOOL), BinaryOp("<", InputRef(1, F64), Constant(10.0), BOOL), BOOL), BOOL)
    got, exp = both(lambda l: ex.try_collect(ex.FilterExecutor(pred, [b, b.slice(0, 0), full], lib=l).execute()), cuda_lib, oracle)
    assert len(got) == len(exp) == 3
    assert_batches_match(got, exp)


# ------------------------------------------------------------------ aggregates
def agg_funcs_for(dtypes):
    out = []
    for i, dt in enumerate(dtypes):
        if dt == I64:
            out += [AggFunc("Sum", [InputRef(i, dt)]), AggFunc("Min", [InputRef(i, dt)]), AggFunc("Max", [InputRef(i, dt)]), AggFunc("Count", [InputRef(i, dt)])]
        elif dt == F64:
            out += [AggFunc("Sum", [InputRef(i, dt)]), AggFunc("Min", [InputRef(i, dt)]), AggFunc("Max", [InputRef(i, dt)])]
        elif dt == I32:
            out += [AggFunc("Min", [InputRef(i, dt)]), AggFunc("Max", [InputRef(i, dt)]), AggFunc("Count", [InputRef(i, dt)]),
                    AggFunc("Sum", [TypeCast(InputRef(i, dt), I64)])]
        else:
            out += [AggFunc("Count", [InputRef(i, dt)])]
    return out


@pytest.mark.parametrize("count_mode", [ffi.COUNT_REFERENCE_OVERWRITE, ffi.COUNT_SQL_ACCUMULATE])
@pytest.mark.parametrize("match_mode", [ffi.MATCH_HASH_ONLY, ffi.MATCH_HASH_AND_KEY])
@pytest.mark.parametrize("cardinality", [3, 13, 4000])
def test_hash_agg_random(cuda_lib, oracle, count_mode, match_mode, cardinality):
    """several nullable batches; 3 groups stay in shared memory, 13 overflow to the HBM table mid-stream, 4000 grow it"""
    rng = np.random.default_rng(cardinality * 7 + count_mode * 3 + match_mode)
    dtypes = [I64, I32, F64, I64, BOOL, I32]
    batches = []
    for n in (700, 0, 2500, 33, 9000):
        b = random_batch(rng, n, dtypes, null_frac=0.1)
        k0 = pa.array(rng.integers(0, cardinality, n).astype(np.int64), mask=rng.random(n) < 0.05)
        k1 = pa.array(rng.integers(0, 2, n).astype(np.int32))
        batches.append(pa.RecordBatch.from_arrays([k0, k1] + b.columns[2:], names=b.schema.names))
    aggs = agg_funcs_for(dtypes)[:14]
    groups = [InputRef(0, I64), InputRef(1, I32)]

    def run(l):
        return ex.try_collect(ex.HashAggExecutor(aggs, groups, batches, lib=l, options=l.options(count_mode=count_mode, match_mode=match_mode)).execute())

    got, exp = both(run, cuda_lib, oracle)
    assert_batches_match(got, exp, rtol=FTOL)


def test_simple_agg_random_and_empty(cuda_lib, oracle):
    rng = np.random.default_rng(77)
    dtypes = [I64, F64, I32, BOOL]
    aggs = agg_funcs_for(dtypes)
    for batches in ([random_batch(rng, 5000, dtypes), random_batch(rng, 1, dtypes), random_batch(rng, 0, dtypes)],
                    [random_batch(rng, 0, dtypes)], [random_batch(rng, 100, dtypes, null_frac=1.0)]):
        for cm in (ffi.COUNT_REFERENCE_OVERWRITE, ffi.COUNT_SQL_ACCUMULATE):
            got, exp = both(lambda l: ex.try_collect(ex.SimpleAggExecutor(aggs, batches, lib=l, options=l.options(count_mode=cm)).execute()), cuda_lib, oracle)
            assert_batches_match(got, exp, rtol=FTOL)
    for lib in (cuda_lib, oracle):  # simple_agg.rs:63 / hash_agg.rs:125 unwrap None when the child yields nothing
        with pytest.raises(ffi.ExecutorError):
            ex.try_collect(ex.SimpleAggExecutor(aggs, [], lib=lib).execute())
        with pytest.raises(ffi.ExecutorError):
            ex.try_collect(ex.HashAggExecutor(aggs, [InputRef(0, I64)], [], lib=lib).execute())


def test_quirk_k2_k3_hash_only_groups(cuda_lib, oracle):
    """combine_hashes is symmetric in two same-typed columns: (0,1) and (1,0) share a row hash, so the reference
    (hash-only identity) merges them and reports the keys of the group's FIRST row; a NULL key keeps hash 0."""
    b = pa.RecordBatch.from_arrays([pa.array([1, 0, 0, 1, None, 5], pa.int64()), pa.array([0, 1, 1, 0, 0, None], pa.int64()),
                                    pa.array([10, 20, 30, 40, 50, 60], pa.int64())], names=["a", "b", "v"])
    aggs, groups = [AggFunc("Sum", [InputRef(2, I64)]), AggFunc("Count", [InputRef(2, I64)])], [InputRef(0, I64), InputRef(1, I64)]
    for mm, n_groups in ((ffi.MATCH_HASH_ONLY, 3), (ffi.MATCH_HASH_AND_KEY, 4)):
        got, exp = both(lambda l: ex.try_collect(ex.HashAggExecutor(aggs, groups, [b], lib=l, options=l.options(match_mode=mm)).execute()), cuda_lib, oracle)
        assert_batches_match(got, exp)
        assert got[0].num_rows == n_groups
    first = rows_of(ex.try_collect(ex.HashAggExecutor(aggs, groups, [b], lib=cuda_lib).execute()))[0]
    assert first == (1, 0, 100, 4)


def test_quirk_k1_count_overwrites(cuda_lib, oracle):
    """count.rs:22 assigns: COUNT = non-NULL count of the LAST batch that touched the group"""
    b1 = pa.RecordBatch.from_arrays([pa.array([1, 1, 2], pa.int64()), pa.array([1, None, 3], pa.int64())], names=["k", "v"])
    b2 = pa.RecordBatch.from_arrays([pa.array([1, 3], pa.int64()), pa.array([7, 8], pa.int64())], names=["k", "v"])
    aggs, groups = [AggFunc("Count", [InputRef(1, I64)]), AggFunc("Sum", [InputRef(1, I64)])], [InputRef(0, I64)]
    got = rows_of(ex.try_collect(ex.HashAggExecutor(aggs, groups, [b1, b2], lib=cuda_lib).execute()))
    exp = rows_of(ex.try_collect(ex.HashAggExecutor(aggs, groups, [b1, b2], lib=oracle).execute()))
    assert got == exp == [(1, 1, 8), (2, 1, 3), (3, 1, 8)]
    sql = cuda_lib.options(count_mode=ffi.COUNT_SQL_ACCUMULATE)
    assert rows_of(ex.try_collect(ex.HashAggExecutor(aggs, groups, [b1, b2], lib=cuda_lib, options=sql).execute())) == [(1, 2, 8), (2, 1, 3), (3, 1, 8)]


# ------------------------------------------------------------------ joins
def _joined_schema(left, right, nullable=True):
    return pa.schema([pa.field("l." + f.name, f.type, nullable) for f in left.schema] + [pa.field("r." + f.name, f.type, nullable) for f in right.schema])


@pytest.mark.parametrize("join_type", ["Inner", "Left", "Right", "Full"])
@pytest.mark.parametrize("match_mode", [ffi.MATCH_HASH_ONLY, ffi.MATCH_HASH_AND_KEY])
@pytest.mark.parametrize("with_filter", [False, True])
def test_hash_join_random(cuda_lib, oracle, join_type, match_mode, with_filter):
    """duplicates on both sides, NULL keys, two key columns, several batches per side, exact output order"""
    rng = np.random.default_rng(hash((join_type, match_mode, with_filter)) % 2**31)
    dtypes = [I64, I32, F64, I64]

    def side(n):
        b = random_batch(rng, n, dtypes, null_frac=0.1, small_ints=True)
        return b

    lefts = [side(300), side(0), side(41)]
    rights = [side(500), side(64), side(0), side(1)]
    cond = ex.JoinCondition([(InputRef(0, I64), InputRef(0, I64)), (InputRef(1, I32), InputRef(1, I32))],
                            BinaryOp(">", InputRef(2, F64), InputRef(6, F64), BOOL) if with_filter else None)
    schema = _joined_schema(lefts[0], rights[0])

    def run(l):
        return ex.try_collect(ex.HashJoinExecutor(lefts, rights, join_type, cond, schema, lib=l, options=l.options(match_mode=match_mode)).execute())

    got, exp = both(run, cuda_lib, oracle)
    assert len(got) == len(exp)
    assert_batches_match(got, exp)


def test_hash_join_heavy_duplicates_and_growth(cuda_lib, oracle):
    """one build key repeated 500x (stable radix-sort path for the CSR), many distinct keys around it"""
    rng = np.random.default_rng(9)
    keys = np.concatenate([np.full(500, 7), rng.integers(0, 20000, 30000)]).astype(np.int64)
    rng.shuffle(keys)
    left = pa.RecordBatch.from_arrays([pa.array(keys), pa.array(np.arange(len(keys), dtype=np.int64))], names=["k", "lrow"])
    pk = np.concatenate([np.full(3, 7), rng.integers(0, 25000, 5000)]).astype(np.int64)
    right = pa.RecordBatch.from_arrays([pa.array(pk), pa.array(np.arange(len(pk), dtype=np.int64))], names=["k", "rrow"])
    cond = ex.JoinCondition([(InputRef(0, I64), InputRef(0, I64))])
    schema = _joined_schema(left, right)
    for jt in ("Inner", "Full"):
        got, exp = both(lambda l: ex.try_collect(ex.HashJoinExecutor([left], [right], jt, cond, schema, lib=l).execute()), cuda_lib, oracle)
        assert_batches_match(got, exp)


def test_join_schema_errors_match(cuda_lib, oracle):
    """RecordBatch::try_new: a non-nullable output field that receives NULLs is an Arrow error"""
    left = pa.RecordBatch.from_arrays([pa.array([1, 2], pa.int64())], names=["a"])
    right = pa.RecordBatch.from_arrays([pa.array([2, 3], pa.int64())], names=["b"])
    schema = _joined_schema(left, right, nullable=False)
    cond = ex.JoinCondition([(InputRef(0, I64), InputRef(0, I64))])
    for lib in (cuda_lib, oracle):
        assert rows_of(ex.try_collect(ex.HashJoinExecutor([left], [right], "Inner", cond, schema, lib=lib).execute())) == [(2, 2)]
        with pytest.raises(ffi.ExecutorError) as err:
            ex.try_collect(ex.HashJoinExecutor([left], [right], "Right", cond, schema, lib=lib).execute())
        assert err.value.code == ffi.ERR_ARROW


# ------------------------------------------------------------------ the benchmark plans, small scale
def _tables(oracle, d):
    return {0: tpch.host_table(oracle, d, tpch.CUSTOMER, columns=tpch.Q3_CUSTOMER_COLUMNS),
            1: tpch.host_table(oracle, d, tpch.ORDERS, columns=tpch.Q3_ORDERS_COLUMNS),
            2: tpch.host_table(oracle, d, tpch.LINEITEM, columns=tpch.Q3_LINEITEM_COLUMNS)}


def _run_plan(lib, plan, schemas, tables, batch_rows=None, **opts):
    p = ExecutorBuilder(lib, lib.options(**opts)).build(plan, schemas)
    for slot, t in tables.items():
        step = batch_rows or max(t.num_rows, 1)
        for off in range(0, max(t.num_rows, 1), step):
            p.push_table(slot, t.slice(off, step))
    out = p.run()
    desc = p.describe()
    p.close()
    return out, desc


@pytest.mark.parametrize("flags_mode", [ffi.TPCH_FLAGS_8GROUP, ffi.TPCH_FLAGS_SPEC])
@pytest.mark.parametrize("modes", [(ffi.COUNT_SQL_ACCUMULATE, ffi.MATCH_HASH_AND_KEY), (ffi.COUNT_REFERENCE_OVERWRITE, ffi.MATCH_HASH_ONLY)])
def test_q1_plan_matches_oracle(cuda_lib, oracle, flags_mode, modes):
    d = tpch.dims(0.05, flags_mode)
    plan, schemas = tpch.q1_plan()
    table = {0: tpch.host_table(oracle, d, tpch.LINEITEM, columns=tpch.Q1_COLUMNS)}
    for fl in (0, ffi.FLAG_NO_FUSION):
        for batch_rows in (None, 65536):
            got, desc = _run_plan(cuda_lib, plan, schemas, table, batch_rows, count_mode=modes[0], match_mode=modes[1], flags=fl)
            exp, _ = _run_plan(oracle, plan, schemas, table, batch_rows, count_mode=modes[0], match_mode=modes[1])
            assert_batches_match(got, exp, rtol=FTOL)
            assert ("fused" in desc) == (fl == 0)


def test_q3_plan_matches_oracle(cuda_lib, oracle):
    d = tpch.dims(0.05)
    plan, schemas = tpch.q3_plan()
    tables = _tables(oracle, d)
    for batch_rows in (None, 100_000):
        got, _ = _run_plan(cuda_lib, plan, schemas, tables, batch_rows, count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
        exp, _ = _run_plan(oracle, plan, schemas, tables, batch_rows, count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
        assert got[0].num_rows > 100
        assert_batches_match(got, exp, rtol=FTOL)


def test_generator_is_bit_identical_on_device(cuda_lib, oracle):
    import torch

    d = tpch.dims(0.02)
    for table in (tpch.CUSTOMER, tpch.ORDERS, tpch.LINEITEM):
        n = tpch.num_rows(oracle, d, table)
        lo, hi = n // 3, n - 5
        host = tpch.host_table(oracle, d, table, lo, hi)
        dev = tpch.device_table(cuda_lib, d, table, lo, hi)
        for c, t in enumerate(dev.tensors):
            want = host.column(c).to_numpy()
            assert np.array_equal(t.cpu().numpy(), want.view(np.int64)), (table, c)


# ------------------------------------------------------------------ BASELINE.json configs[1] / configs[2] at their full size, vs the oracle
def _assert_tables_equal(got, exp, rtol):
    """Vectorised: every integer column and the row order bit-exact, Float64 within rtol (north star: 1e-6)."""
    got, exp = pa.Table.from_batches(got).combine_chunks(), pa.Table.from_batches(exp).combine_chunks()
    assert got.schema.names == exp.schema.names and got.num_rows == exp.num_rows, (got.schema, exp.schema, got.num_rows, exp.num_rows)
    for name in got.schema.names:
        g, w = got.column(name), exp.column(name)
        assert g.null_count == w.null_count == 0, name
        g, w = g.to_numpy(), w.to_numpy()
        if g.dtype.kind == "f":
            bad = np.abs(g - w) > rtol * np.abs(w)
            assert not bad.any(), (name, int(bad.sum()), g[bad][:5], w[bad][:5])
        else:
            assert np.array_equal(g, w), (name, int((g != w).sum()))


def test_q3_sf10_bit_exact_vs_oracle(cuda_lib, oracle):
    """BASELINE.json configs[2]: "TPC-H Q3 SF10 ... bit-exact row check" — 76.5 M input rows, ~113 k result groups: keys,
    group order (first appearance, hash_agg.rs:98,134) exact; SUM(float) <= 1e-6 relative.  The oracle runs the reference's
    per-batch algorithm on one batch per table (~5 s)."""
    d = tpch.dims(10)
    plan, schemas = tpch.q3_plan()
    tables = _tables(oracle, d)
    opts = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    got, desc = _run_plan(cuda_lib, plan, schemas, tables, None, **opts)
    exp, _ = _run_plan(oracle, plan, schemas, tables, None, **opts)
    assert sum(b.num_rows for b in exp) > 50_000
    _assert_tables_equal(got, exp, 1e-6)
    # the whole query (ORDER BY revenue desc, o_orderdate LIMIT 10 on the device) against the oracle's own Order/Limit
    full, _ = tpch.q3_full_plan()
    got10, _ = _run_plan(cuda_lib, full, schemas, tables, None, **opts)
    exp10, _ = _run_plan(oracle, full, schemas, tables, None, **opts)
    assert sum(b.num_rows for b in got10) == 10
    _assert_tables_equal(got10, exp10, 1e-6)


@pytest.mark.parametrize("batch_rows", [None, 1 << 22])
def test_q1_sf10_bit_exact_vs_oracle(cuda_lib, oracle, batch_rows):
    """BASELINE.json configs[1]: Q1' SF10 (60 M rows): 8 groups, COUNT and the int64 SUM exact, Float64 sums <= 1e-6."""
    d = tpch.dims(10)
    plan, schemas = tpch.q1_plan()
    table = {0: tpch.host_table(oracle, d, tpch.LINEITEM, columns=tpch.Q1_COLUMNS)}
    opts = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    got, _ = _run_plan(cuda_lib, plan, schemas, table, batch_rows, **opts)
    exp, _ = _run_plan(oracle, plan, schemas, table, 1 << 22, **opts)
    assert exp[0].num_rows == 8
    _assert_tables_equal(got, exp, 1e-6)


# ------------------------------------------------------------------ full size: properties that need no oracle pass
def test_q1_full_size_properties(cuda_lib):
    """BASELINE config 2 size (SF10, ~60 M rows, resident in HBM): integer checksums and linearity"""
    import torch

    d = tpch.dims(10)
    n = tpch.num_rows(cuda_lib, d, tpch.LINEITEM)
    plan, schemas = tpch.q1_plan()
    opts = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)

    def run(lo, hi):
        t = tpch.device_table(cuda_lib, d, tpch.LINEITEM, lo, hi, columns=tpch.Q1_COLUMNS)
        p = ExecutorBuilder(cuda_lib, cuda_lib.options(**opts)).build(plan, schemas)
        p.push_table_device(0, t)
        out = pa.Table.from_batches(p.run()).to_pydict()
        p.close()
        return {(a, b): [out[c][i] for c in list(out)[2:]] for i, (a, b) in enumerate(zip(out["l_returnflag"], out["l_linestatus"]))}, t

    whole, table = run(0, n)
    assert len(whole) == 8
    ship = table.tensors[tpch.Q1_COLUMNS.index(7)]
    qty = table.tensors[tpch.Q1_COLUMNS.index(8)]
    keep = ship <= tpch.Q1_SHIPDATE_MAX
    assert sum(v[5] for v in whole.values()) == int(keep.sum().item())            # count of counts
    assert sum(v[4] for v in whole.values()) == int(qty[keep].sum().item())        # checksum of the integer sums
    for v in whole.values():
        assert v[0] == float(v[4])   # SUM(l_quantity) over doubles of small integers is exact and equals the int64 sum
    del table
    torch.cuda.empty_cache()
    a, _ = run(0, n // 3)
    b, _ = run(n // 3, n)
    for k, v in whole.items():       # linearity: two shards add up (ints exactly, float sums to rounding)
        assert v[4] == a[k][4] + b[k][4] and v[5] == a[k][5] + b[k][5]
        for j in range(4):
            assert abs(v[j] - (a[k][j] + b[k][j])) <= 1e-9 * abs(v[j])


def test_device_resident_partial_exchange(cuda_lib, oracle):
    """The NCCL fast path's building blocks on one GPU: two shards aggregated separately, their packed partial
    tables (device memory) concatenated like an all-gather would, folded and finalised == whole-table result."""
    import torch

    d = tpch.dims(0.05)
    plan, schemas = tpch.q1_plan()
    opts = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    whole = tpch.host_table(oracle, d, tpch.LINEITEM, columns=tpch.Q1_COLUMNS)
    n = whole.num_rows
    cut = n // 3
    cap = 64
    bufs, plans = [], []
    for lo, hi in ((0, cut), (cut, n)):
        p = ExecutorBuilder(cuda_lib, cuda_lib.options(**opts)).build(plan, schemas)
        p.push_table(0, whole.slice(lo, hi - lo))
        cuda_lib.check(cuda_lib.plan_execute_partial(p.handle, lo))
        words = C.c_int32(0)
        cuda_lib.check(cuda_lib.plan_partials_row_words(p.handle, C.byref(words)))
        buf = torch.empty((cap + 1) * words.value, dtype=torch.int64, device="cuda")
        cuda_lib.check(cuda_lib.plan_export_partials_device(p.handle, C.c_void_p(buf.data_ptr()), cap))
        torch.cuda.synchronize()
        assert int(buf[0].item()) == 8
        bufs.append(buf)
        plans.append(p)
    gathered = torch.cat(bufs)
    p = plans[0]
    cuda_lib.check(cuda_lib.plan_clear_partials(p.handle))
    cuda_lib.check(cuda_lib.plan_merge_partials_device(p.handle, C.c_void_p(gathered.data_ptr()), 2, cap))
    cuda_lib.check(cuda_lib.plan_finish_partial(p.handle))
    got = p.collect()
    exp, _ = _run_plan(oracle, plan, schemas, {0: whole}, None, **opts)
    assert_batches_match(got, exp, rtol=FTOL)
    # the host (Arrow) form of the same exchange
    from sqlrs_b200.host import distributed as sqdist

    q = plans[1]
    part = sqdist._export_partials(q)
    assert part.num_rows == 8 and part.schema.names[:3] == ["hash", "min_row", "knull"]
    cuda_lib.check(cuda_lib.plan_clear_partials(q.handle))
    for piece in sqdist.partition_by_owner(part, 3):
        sqdist._merge_partials(q, piece)
    cuda_lib.check(cuda_lib.plan_finish_partial(q.handle))
    shard_only, _ = _run_plan(oracle, plan, schemas, {0: whole.slice(cut)}, None, **opts)
    assert_batches_match(q.collect(), shard_only, rtol=FTOL)
    for x in plans:
        x.close()


@pytest.mark.parametrize("grouped", [True, False])
def test_distinct_aggregates_in_the_partial_final_split(cuda_lib, oracle, grouped):
    """DISTINCT aggregates keep one table per set next to the plain aggregates' table: two shards aggregated separately, every
    table exported (device form for one shard, host Arrow form for the other), merged into a third plan == whole-table result"""
    import torch
    from sqlrs_b200.host import distributed as sqdist
    from sqlrs_b200.host.plan import PhysicalHashAgg, PhysicalSimpleAgg, PhysicalTableScan

    rng = np.random.default_rng(77)
    n = 20_000
    schema = pa.schema([pa.field("k", pa.int64()), pa.field("x", pa.int64()), pa.field("v", pa.int64())])
    k = rng.integers(0, 300, n)
    x = pa.array(rng.integers(0, 40, n), mask=rng.random(n) < 0.05)
    whole = pa.RecordBatch.from_arrays([pa.array(k), x, pa.array(rng.integers(-1000, 1000, n))], schema=schema)
    K, X, V = InputRef(0, I64), InputRef(1, I64), InputRef(2, I64)
    aggs = [AggFunc("Count", [X], distinct=True), AggFunc("Sum", [V]), AggFunc("Sum", [X], distinct=True), AggFunc("Count", [V])]
    plan = PhysicalHashAgg(aggs, [K], PhysicalTableScan(0)) if grouped else PhysicalSimpleAgg(aggs, PhysicalTableScan(0))
    opts = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    exp, _ = _run_plan(oracle, plan, {0: schema}, {0: whole}, None, **opts)
    cut = n // 3
    plans = []
    for lo, hi in ((0, cut), (cut, n), (0, 1)):  # the third plan is the merge target (its own one-row partial state is cleared)
        p = ExecutorBuilder(cuda_lib, cuda_lib.options(**opts)).build(plan, {0: schema})
        p.push_table(0, whole.slice(lo, hi - lo))
        cuda_lib.check(cuda_lib.plan_execute_partial(p.handle, lo))
        plans.append(p)
    a, b, target = plans
    tables = C.c_int32(0)
    cuda_lib.check(cuda_lib.plan_partials_tables(target.handle, C.byref(tables)))
    assert tables.value == 3  # plain aggregates + two sets
    for t in range(tables.value):
        for p in plans:
            cuda_lib.check(cuda_lib.plan_select_partials_table(p.handle, t))
        cuda_lib.check(cuda_lib.plan_clear_partials(target.handle))
        words = C.c_int32(0)
        cuda_lib.check(cuda_lib.plan_partials_row_words(a.handle, C.byref(words)))
        cap = n
        buf = torch.empty((cap + 1) * words.value, dtype=torch.int64, device="cuda")
        cuda_lib.check(cuda_lib.plan_export_partials_device(a.handle, C.c_void_p(buf.data_ptr()), cap))
        torch.cuda.synchronize()
        cuda_lib.check(cuda_lib.plan_merge_partials_device(target.handle, C.c_void_p(buf.data_ptr()), 1, cap))
        for piece in sqdist.partition_by_owner(sqdist._export_partials(b), 2):
            sqdist._merge_partials(target, piece)
    cuda_lib.check(cuda_lib.plan_finish_partial(target.handle))
    assert_batches_match(target.collect(), exp)
    with pytest.raises(ffi.ExecutorError):
        cuda_lib.check(cuda_lib.plan_execute_partial(a.handle, 0))
        cuda_lib.check(cuda_lib.plan_select_partials_table(a.handle, 3))
    for p in plans:
        p.close()


def test_partial_run_defers_the_tier_check_and_recovers(cuda_lib, oracle):
    """execute_partial leaves the first sq_agg_small launch unchecked (no synchronisation before the exchange).  With more groups
    than that tier holds, the packed device export carries an impossible count (every rank of an exchange sees it), and any
    call that needs the real state re-runs the partial aggregation the checked way; later runs of the plan do not defer."""
    import torch
    from sqlrs_b200.host import distributed as sqdist
    from sqlrs_b200.host.plan import PhysicalHashAgg, PhysicalTableScan

    rng = np.random.default_rng(5)
    n, groups = 300_000, 2000
    schema = pa.schema([pa.field("k", pa.int64()), pa.field("v", pa.int64())])
    table = pa.RecordBatch.from_arrays([pa.array(rng.integers(0, groups, n)), pa.array(rng.integers(-100, 100, n))], schema=schema)
    plan = PhysicalHashAgg([AggFunc("Sum", [InputRef(1, I64)]), AggFunc("Count", [InputRef(1, I64)])], [InputRef(0, I64)], PhysicalTableScan(0))
    opts = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    exp, _ = _run_plan(oracle, plan, {0: schema}, {0: table}, None, **opts)
    p = ExecutorBuilder(cuda_lib, cuda_lib.options(**opts)).build(plan, {0: schema})
    p.push_table(0, table)
    words = C.c_int32(0)
    cap = 1 << 16
    for attempt in range(2):
        cuda_lib.check(cuda_lib.plan_execute_partial(p.handle, 0))
        cuda_lib.check(cuda_lib.plan_partials_row_words(p.handle, C.byref(words)))
        buf = torch.zeros((cap + 1) * words.value, dtype=torch.int64, device="cuda")
        cuda_lib.check(cuda_lib.plan_export_partials_device(p.handle, C.c_void_p(buf.data_ptr()), cap))
        torch.cuda.synchronize()
        header = int(buf[0].item())
        if attempt == 0:
            assert header >= 1 << 40, header  # unchecked and overflowed: marked
        else:
            assert header == groups  # the operator remembered: checked run, the real group table
        part = sqdist._export_partials(p)  # needs the real state: settles, re-runs if the deferred launch had overflowed
        assert part.num_rows == groups
        cuda_lib.check(cuda_lib.plan_clear_partials(p.handle))
        sqdist._merge_partials(p, part)
        cuda_lib.check(cuda_lib.plan_finish_partial(p.handle))
        assert_batches_match(p.collect(), exp)
    p.close()


def test_partial_merge_hash_only_first_rows_keys_win(cuda_lib, oracle):
    """Quirk K2 across partial tables: (i, j) and (j, i) share a row hash; under hash-only identity the merged group
    reports the keys of the globally FIRST row, whichever partial supplied it and whatever order the merge saw them in."""
    import torch
    from sqlrs_b200.host.plan import PhysicalHashAgg, PhysicalTableScan

    m = 3000
    i = np.arange(m, dtype=np.int64)
    schema = pa.schema([pa.field("a", pa.int64()), pa.field("b", pa.int64()), pa.field("v", pa.int64())])
    # shard 0 (global rows 0..m): even groups as (lo, hi), odd groups as (hi, lo); shard 1 the mirror image
    a0 = pa.array(np.where(i % 2 == 0, i, i + 100_000))
    b0 = pa.array(np.where(i % 2 == 0, i + 100_000, i))
    shard0 = pa.RecordBatch.from_arrays([a0, b0, pa.array(i + 1)], schema=schema)
    shard1 = pa.RecordBatch.from_arrays([b0, a0, pa.array(2 * i + 1)], schema=schema)
    plan = PhysicalHashAgg([AggFunc("Sum", [InputRef(2, I64)])], [InputRef(0, I64), InputRef(1, I64)], PhysicalTableScan(0))
    opts = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_ONLY)
    whole = pa.Table.from_batches([shard0, shard1]).combine_chunks().to_batches()[0]
    exp, _ = _run_plan(oracle, plan, {0: schema}, {0: whole}, None, **opts)
    assert exp[0].num_rows == m
    cap = 4096
    for order in ((0, 1), (1, 0)):
        bufs, plans = [], []
        for shard, base in ((shard0, 0), (shard1, m)):
            p = ExecutorBuilder(cuda_lib, cuda_lib.options(**opts)).build(plan, {0: schema})
            p.push_table(0, shard)
            cuda_lib.check(cuda_lib.plan_execute_partial(p.handle, base))
            words = C.c_int32(0)
            cuda_lib.check(cuda_lib.plan_partials_row_words(p.handle, C.byref(words)))
            buf = torch.empty((cap + 1) * words.value, dtype=torch.int64, device="cuda")
            cuda_lib.check(cuda_lib.plan_export_partials_device(p.handle, C.c_void_p(buf.data_ptr()), cap))
            torch.cuda.synchronize()
            bufs.append(buf)
            plans.append(p)
        gathered = torch.cat([bufs[order[0]], bufs[order[1]]])
        p = plans[0]
        cuda_lib.check(cuda_lib.plan_clear_partials(p.handle))
        cuda_lib.check(cuda_lib.plan_merge_partials_device(p.handle, C.c_void_p(gathered.data_ptr()), 2, cap))
        cuda_lib.check(cuda_lib.plan_finish_partial(p.handle))
        assert_batches_match(p.collect(), exp)
        for x in plans:
            x.close()


def test_plan_can_be_executed_repeatedly(cuda_lib, oracle):
    """bench.py reuses one plan: operator state must reset between runs (and survive a different table)"""
    plan, schemas = tpch.q1_plan()
    opts = dict(count_mode=ffi.COUNT_REFERENCE_OVERWRITE, match_mode=ffi.MATCH_HASH_ONLY)
    p = ExecutorBuilder(cuda_lib, cuda_lib.options(**opts)).build(plan, schemas)
    for sf, mode in ((0.01, ffi.TPCH_FLAGS_8GROUP), (0.02, ffi.TPCH_FLAGS_SPEC), (0.01, ffi.TPCH_FLAGS_8GROUP)):
        t = tpch.host_table(oracle, tpch.dims(sf, mode), tpch.LINEITEM, columns=tpch.Q1_COLUMNS)
        p.push_table(0, t.slice(0, 5000))
        p.push_table(0, t.slice(5000))
        got = p.run()
        p.reset()
        exp, _ = _run_plan(oracle, plan, schemas, {0: t}, None, **opts)
        q = ExecutorBuilder(oracle, oracle.options(**opts)).build(plan, schemas)
        q.push_table(0, t.slice(0, 5000))
        q.push_table(0, t.slice(5000))
        assert_batches_match(got, q.run(), rtol=FTOL)
    p.close()


# ------------------------------------------------------------------ plan executor: fused side Filters, column pruning
def _random_join_plan(rng, join_type, with_join_filter, root_agg):
    """HashAgg?(HashJoin(Filter(scan 0), Filter(scan 1))) over two random tables."""
    from sqlrs_b200.host.plan import PhysicalFilter, PhysicalHashAgg, PhysicalHashJoin, PhysicalTableScan

    dtypes = [I64, I32, F64, I64]
    left = random_batch(rng, 400, dtypes, null_frac=0.1, small_ints=True, names=["la", "lb", "lc", "ld"])
    right = random_batch(rng, 700, dtypes, null_frac=0.1, small_ints=True, names=["ra", "rb", "rc", "rd"])
    schema = pa.schema([pa.field("l." + f.name, f.type, True) for f in left.schema] + [pa.field("r." + f.name, f.type, True) for f in right.schema])
    lpred = bind_binary_op(InputRef(3, I64), ">", Constant(-4))
    rpred = BinaryOp("OR", bind_binary_op(InputRef(3, I64), "<", Constant(5)), BinaryOp(">", InputRef(2, F64), Constant(50.0), BOOL), BOOL)
    cond = ex.JoinCondition([(InputRef(0, I64), InputRef(0, I64))], BinaryOp("<", InputRef(2, F64), InputRef(6, F64), BOOL) if with_join_filter else None)
    join = PhysicalHashJoin(PhysicalFilter(lpred, PhysicalTableScan(0)), PhysicalFilter(rpred, PhysicalTableScan(1)), join_type, cond, schema)
    root = join
    if root_agg:
        root = PhysicalHashAgg([AggFunc("Sum", [InputRef(6, F64)]), AggFunc("Count", [InputRef(1, I32)]), AggFunc("Max", [InputRef(7, I64)])],
                               [InputRef(4, I64), InputRef(1, I32)], join)
    return root, {0: left.schema, 1: right.schema}, {0: left, 1: right}


@pytest.mark.parametrize("join_type", ["Inner", "Left", "Right", "Full"])
@pytest.mark.parametrize("with_join_filter", [False, True])
@pytest.mark.parametrize("root_agg", [False, True])
def test_plan_join_with_side_filters(cuda_lib, oracle, join_type, with_join_filter, root_agg):
    """Filters below a join run inside the join's key kernels and unread columns are not gathered when the plan is
    fused; the result must equal operator-at-a-time execution and the oracle, row order included."""
    rng = np.random.default_rng(hash((join_type, with_join_filter, root_agg)) % 2**31)
    root, schemas, tables = _random_join_plan(rng, join_type, with_join_filter, root_agg)
    opts = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    exp, _ = _run_plan(oracle, root, schemas, tables, None, **opts)
    for flags in (0, ffi.FLAG_NO_FUSION):
        for batch_rows in (None, 150):
            got, desc = _run_plan(cuda_lib, root, schemas, tables, batch_rows, flags=flags, **opts)
            if batch_rows is None:
                assert_batches_match(got, exp, rtol=FTOL)
            else:
                exp_b, _ = _run_plan(oracle, root, schemas, tables, batch_rows, **opts)
                assert_batches_match(got, exp_b, rtol=FTOL)
            assert ("fused" in desc) == (flags == 0)


@pytest.mark.parametrize("key_type", [I32, F64, BOOL])
def test_group_by_and_join_on_other_key_types(cuda_lib, oracle, key_type):
    """Int32 / Float64 (incl. -0.0, NaN, inf) / Boolean keys: hash_one of the reference's other primitive types"""
    rng = np.random.default_rng(int(key_type) + 40)
    n = 3000
    if key_type == F64:
        pool = np.array([0.0, -0.0, 1.5, -1.5, np.inf, -np.inf, np.nan, 1e300, 2.5, 3.5])
        keys = pa.array(pool[rng.integers(0, len(pool), n)], mask=rng.random(n) < 0.05)
    elif key_type == BOOL:
        keys = pa.array(rng.integers(0, 2, n).astype(bool), mask=rng.random(n) < 0.1)
    else:
        keys = pa.array(rng.integers(-3, 4, n).astype(np.int32), mask=rng.random(n) < 0.1)
    vals = pa.array(rng.integers(-100, 100, n).astype(np.int64))
    b = pa.RecordBatch.from_arrays([keys, vals], names=["k", "v"])
    aggs, groups = [AggFunc("Sum", [InputRef(1, I64)]), AggFunc("Count", [InputRef(0, key_type)])], [InputRef(0, key_type)]
    for mm in (ffi.MATCH_HASH_ONLY, ffi.MATCH_HASH_AND_KEY):
        got, exp = both(lambda l: ex.try_collect(ex.HashAggExecutor(aggs, groups, [b.slice(0, 1000), b.slice(1000)], lib=l,
                                                                    options=l.options(match_mode=mm, count_mode=ffi.COUNT_SQL_ACCUMULATE)).execute()), cuda_lib, oracle)
        g, e = rows_of(got), rows_of(exp)
        assert len(g) == len(e)
        for x, y in zip(g, e):
            same_key = (x[0] == y[0]) or (x[0] is not None and y[0] is not None and x[0] != x[0] and y[0] != y[0])
            assert same_key and x[1:] == y[1:], (x, y)
        small = b.slice(0, 60)
        schema = pa.schema([pa.field("l.k", keys.type), pa.field("l.v", pa.int64()), pa.field("r.k", keys.type), pa.field("r.v", pa.int64())])
        cond = ex.JoinCondition([(InputRef(0, key_type), InputRef(0, key_type))])
        gj, ej = both(lambda l: ex.try_collect(ex.HashJoinExecutor([small], [b.slice(60, 200)], "Full", cond, schema, lib=l,
                                                                   options=l.options(match_mode=mm)).execute()), cuda_lib, oracle)
        assert [(r[1], r[3]) for r in rows_of(gj)] == [(r[1], r[3]) for r in rows_of(ej)]


def test_many_groups_with_growth_across_batches(cuda_lib, oracle):
    """cardinality climbs batch by batch: the HBM table is re-hashed several times while hash-only key fix-ups are pending"""
    rng = np.random.default_rng(321)
    batches = []
    for i, n in enumerate((500, 5000, 60000, 250000)):
        k = pa.array(rng.integers(0, 40 * (i + 1) ** 4, n).astype(np.int64))
        k2 = pa.array(rng.integers(0, 3, n).astype(np.int64))
        v = pa.array(rng.integers(-1000, 1000, n).astype(np.int64))
        f = pa.array(np.round(rng.normal(0, 10, n), 3), mask=rng.random(n) < 0.2)
        batches.append(pa.RecordBatch.from_arrays([k, k2, v, f], names=["k", "k2", "v", "f"]))
    aggs = [AggFunc("Sum", [InputRef(2, I64)]), AggFunc("Count", [InputRef(3, F64)]), AggFunc("Min", [InputRef(3, F64)]), AggFunc("Sum", [InputRef(3, F64)])]
    groups = [InputRef(0, I64), InputRef(1, I64)]
    for mm in (ffi.MATCH_HASH_ONLY, ffi.MATCH_HASH_AND_KEY):
        got, exp = both(lambda l: ex.try_collect(ex.HashAggExecutor(aggs, groups, batches, lib=l, options=l.options(match_mode=mm)).execute()), cuda_lib, oracle)
        assert got[0].num_rows > 10000
        assert_batches_match(got, exp, rtol=FTOL)


# ------------------------------------------------------------------ Order / Limit / Project
def _special_floats(rng, n):
    v = np.round(rng.normal(0, 3, n), 0)  # many ties
    for k, x in enumerate((np.nan, -np.nan, np.inf, -np.inf, 0.0, -0.0, 5e-324, -1e308)):
        v[rng.integers(0, n, 3)] = x
    return v


@pytest.mark.parametrize("seed", range(8))
def test_order_random(cuda_lib, oracle, seed):
    """ties, NULLs, NaN / -0.0 / inf, every key type, 1..3 sort keys in both directions, several input batches: the
    output row order must equal the oracle's exactly (both keep ties in input order)"""
    rng = np.random.default_rng(1000 + seed)
    dtypes = [I64, F64, I32, BOOL, I64]
    batches = []
    for n in (0, 700, 1, 3000)[: 2 + seed % 3]:
        b = random_batch(rng, n, dtypes, null_frac=0.2 if seed % 2 else 0.0, small_ints=True)
        if n:
            mask = None if seed % 2 == 0 else rng.random(n) < 0.2
            b = b.set_column(1, "c1", pa.array(_special_floats(rng, n), mask=mask))
        batches.append(b)
    n_keys = 1 + seed % 3
    cols = rng.permutation(4)[:n_keys]
    order_by = [ex.BoundOrderBy(InputRef(int(c), dtypes[int(c)]), asc=bool(rng.integers(0, 2))) for c in cols]
    if seed == 5:  # an expression as the sort key
        order_by = [ex.BoundOrderBy(bind_binary_op(InputRef(0, I64), "*", InputRef(4, I64)), asc=False)]
    got, exp = both(lambda l: ex.try_collect(ex.OrderExecutor(order_by, batches, lib=l).execute()), cuda_lib, oracle)
    assert len(got) == len(exp) == 1 and got[0].num_rows == sum(b.num_rows for b in batches)
    assert_batches_match(got, exp)


def test_order_single_key_descending_reverses_null_run(cuda_lib, oracle):
    """arrow's single-column sort_to_indices reverses the run of NULL rows when descending"""
    b = pa.RecordBatch.from_arrays([pa.array([3, None, 1, None, 2, None], pa.int64()), pa.array([0, 1, 2, 3, 4, 5], pa.int64())], names=["k", "row"])
    for lib_ in (cuda_lib, oracle):
        out = ex.try_collect(ex.OrderExecutor([ex.BoundOrderBy(InputRef(0, I64), asc=False)], [b], lib=lib_).execute())
        assert [r[1] for r in rows_of(out)] == [5, 3, 1, 0, 4, 2]
        out = ex.try_collect(ex.OrderExecutor([ex.BoundOrderBy(InputRef(0, I64), asc=True)], [b], lib=lib_).execute())
        assert [r[1] for r in rows_of(out)] == [1, 3, 5, 2, 4, 0]


@pytest.mark.parametrize("limit,offset", [(5, None), (None, 7), (1000, 3), (10, 995), (0, 0), (4000, 0), (300, 700), (1, 2999)])
def test_limit_random_batches(cuda_lib, oracle, limit, offset):
    rng = np.random.default_rng(7)
    batches = [random_batch(rng, n, [I64, F64, BOOL, I32]) for n in (400, 0, 600, 33, 2000)]
    def run(l):
        try:
            return ex.try_collect(ex.LimitExecutor(limit, offset, batches, lib=l).execute())
        except ffi.ExecutorError as e:  # OFFSET without LIMIT over several batches underflows at limit.rs:58 (the reference panics)
            return e.code

    got, exp = both(run, cuda_lib, oracle)
    if isinstance(exp, int):
        assert got == exp == ffi.ERR_INTERNAL and limit is None
        return
    assert [b.num_rows for b in got] == [b.num_rows for b in exp]
    assert_batches_match(got, exp)


@pytest.mark.parametrize("seed", range(4))
def test_project_random(cuda_lib, oracle, seed):
    rng = np.random.default_rng(50 + seed)
    dtypes = [I64, F64, I32, BOOL, I64, F64]
    batches = [random_batch(rng, n, dtypes, small_ints=True) for n in (257, 0, 4000)]
    exprs = [InputRef(3, BOOL), random_expr(rng, dtypes, I64, 3), random_expr(rng, dtypes, F64, 2), InputRef(1, F64), random_expr(rng, dtypes, BOOL, 3)]
    got, exp = both(lambda l: ex.try_collect(ex.ProjectExecutor(exprs, batches, lib=l).execute()), cuda_lib, oracle)
    assert_batches_match(got, exp, rtol=FTOL)


def test_aggregate_finalised_on_device(cuda_lib, oracle):
    """> 1024 groups take the device finalisation path (packed rows -> typed columns + validity by one kernel): every
    accumulator kind, nullable arguments, a nullable key, Float64 MIN/MAX, both COUNT modes, first-appearance order"""
    rng = np.random.default_rng(99)
    n = 120_000
    k = pa.array(rng.integers(0, 5000, n).astype(np.int64), mask=rng.random(n) < 0.01)
    k2 = pa.array(rng.integers(0, 2, n).astype(bool))
    v = pa.array(rng.integers(-1000, 1000, n).astype(np.int64), mask=rng.random(n) < 0.5)
    f = pa.array(np.round(rng.normal(0, 10, n), 3), mask=rng.random(n) < 0.5)
    w = pa.array(rng.integers(-100, 100, n).astype(np.int32))
    b = pa.RecordBatch.from_arrays([k, k2, v, f, w], names=["k", "k2", "v", "f", "w"])
    batches = [b.slice(0, 50_000), b.slice(50_000)]
    aggs = [AggFunc("Sum", [InputRef(2, I64)]), AggFunc("Count", [InputRef(3, F64)]), AggFunc("Min", [InputRef(3, F64)]), AggFunc("Max", [InputRef(3, F64)]),
            AggFunc("Sum", [InputRef(3, F64)]), AggFunc("Max", [InputRef(4, I32)]), AggFunc("Count", [InputRef(0, I64)])]
    groups = [InputRef(0, I64), InputRef(1, BOOL)]
    for cm in (ffi.COUNT_REFERENCE_OVERWRITE, ffi.COUNT_SQL_ACCUMULATE):
        opts = dict(count_mode=cm, match_mode=ffi.MATCH_HASH_AND_KEY)
        got, exp = both(lambda l: ex.try_collect(ex.HashAggExecutor(aggs, groups, batches, lib=l, options=l.options(**opts)).execute()), cuda_lib, oracle)
        assert got[0].num_rows > 5000
        assert_batches_match(got, exp, rtol=FTOL)


def test_q3_full_plan_with_tail_on_device(cuda_lib, oracle):
    """Q3' including ORDER BY revenue desc, o_orderdate LIMIT 10 and the select list: fused (top-k gather), operator at a
    time, and the oracle agree row for row"""
    d = tpch.dims(0.05)
    plan, schemas = tpch.q3_full_plan()
    tables = _tables(oracle, d)
    mode = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    exp, _ = _run_plan(oracle, plan, schemas, tables, None, **mode)
    assert exp[0].num_rows == 10 and exp[0].schema.names == ["lineitem.l_orderkey", "revenue", "orders.o_orderdate", "orders.o_shippriority"]
    for fl in (0, ffi.FLAG_NO_FUSION):
        for batch_rows in (None, 100_000):
            got, desc = _run_plan(cuda_lib, plan, schemas, tables, batch_rows, flags=fl, **mode)
            expb, _ = _run_plan(oracle, plan, schemas, tables, batch_rows, **mode)
            assert_batches_match(got, expb, rtol=FTOL)
            assert ("Order" in desc and "Limit" in desc) or desc.startswith("oracle")


def test_q1_full_plan_with_tail_on_device(cuda_lib, oracle):
    d = tpch.dims(0.05)
    plan, schemas = tpch.q1_full_plan()
    table = {0: tpch.host_table(oracle, d, tpch.LINEITEM, columns=tpch.Q1_COLUMNS)}
    mode = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    got, _ = _run_plan(cuda_lib, plan, schemas, table, 65536, **mode)
    exp, _ = _run_plan(oracle, plan, schemas, table, 65536, **mode)
    assert [r[:2] for r in rows_of(got)] == sorted(r[:2] for r in rows_of(got)) and got[0].num_rows == 8
    assert_batches_match(got, exp, rtol=FTOL)


def test_plan_order_limit_over_join_and_filter(cuda_lib, oracle):
    """tail operators over non-aggregate children: Limit(Order(Filter(scan))) and Limit(Project(HashJoin)), several batches"""
    from sqlrs_b200.host.plan import PhysicalFilter, PhysicalHashJoin, PhysicalLimit, PhysicalOrder, PhysicalProject, PhysicalTableScan

    rng = np.random.default_rng(4242)
    t = random_batch(rng, 5000, [I64, F64, I64], small_ints=True, names=["a", "b", "c"])
    u = random_batch(rng, 300, [I64, I64], small_ints=True, names=["x", "y"])
    schemas = {0: t.schema, 1: u.schema}
    tables = {0: t, 1: u}
    filt = PhysicalFilter(bind_binary_op(InputRef(0, I64), ">", Constant(-3)), PhysicalTableScan(0))
    p1 = PhysicalLimit(40, 10, PhysicalOrder([ex.BoundOrderBy(InputRef(1, F64), False), ex.BoundOrderBy(InputRef(2, I64), True)], filt))
    jschema = pa.schema([pa.field("u.x", pa.int64()), pa.field("u.y", pa.int64()), pa.field("t.a", pa.int64()), pa.field("t.b", pa.float64()), pa.field("t.c", pa.int64())])
    join = PhysicalHashJoin(PhysicalTableScan(1), PhysicalTableScan(0), "Inner", ex.JoinCondition([(InputRef(0, I64), InputRef(0, I64))]), jschema)
    p2 = PhysicalLimit(500, 100, PhysicalProject([bind_binary_op(InputRef(1, I64), "+", InputRef(4, I64)), InputRef(3, F64)], join))
    for plan in (p1, p2):
        for batch_rows in (None, 1000):
            for fl in (0, ffi.FLAG_NO_FUSION):
                got, _ = _run_plan(cuda_lib, plan, schemas, tables, batch_rows, flags=fl, match_mode=ffi.MATCH_HASH_AND_KEY)
                exp, _ = _run_plan(oracle, plan, schemas, tables, batch_rows, match_mode=ffi.MATCH_HASH_AND_KEY)
                assert sum(b.num_rows for b in got) > 0
                assert_batches_match(got, exp, rtol=FTOL)


def test_q3_tma_staged_probe_variant(cuda_lib, oracle, monkeypatch):
    """the opt-in TMA variant of the fused probe + aggregate kernel (cp.async.bulk tiles + mbarrier ring, SQLRS_B200_TMA=1)
    gives the same groups; a sliced (16-byte misaligned) probe table silently takes the register-staged kernel"""
    monkeypatch.setenv("SQLRS_B200_TMA", "1")
    d = tpch.dims(0.05)
    plan, schemas = tpch.q3_plan()
    tables = _tables(oracle, d)
    mode = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    exp, _ = _run_plan(oracle, plan, schemas, tables, None, **mode)
    got, desc = _run_plan(cuda_lib, plan, schemas, tables, None, **mode)
    assert_batches_match(got, exp, rtol=FTOL)
    assert "sq_joinagg_tma_kernel" in desc or desc.startswith("oracle")
    sliced = dict(tables)
    sliced[2] = tables[2].slice(1)  # 8-byte offset into every column buffer
    exp2, _ = _run_plan(oracle, plan, schemas, sliced, None, **mode)
    got2, desc2 = _run_plan(cuda_lib, plan, schemas, sliced, None, **mode)
    assert_batches_match(got2, exp2, rtol=FTOL)


@pytest.mark.parametrize("match_mode", [ffi.MATCH_HASH_ONLY, ffi.MATCH_HASH_AND_KEY])
def test_distinct_aggregates_random(cuda_lib, oracle, match_mode):
    """COUNT(DISTINCT x) / SUM(DISTINCT x) mixed with plain aggregates, NULLs in arguments and keys, several batches, grouped and
    ungrouped, Int64 / Int32 / Boolean arguments (integer sums are order-independent, so results are exact); `select distinct`
    = a group-by without aggregates"""
    rng = np.random.default_rng(77)
    batches = [random_batch(rng, n, [I64, I64, I32, BOOL, I64], null_frac=0.2, small_ints=True, names=["k", "x", "y", "b", "v"]) for n in (3000, 0, 5000)]
    k, x, y, b, v = (InputRef(i, t) for i, t in enumerate([I64, I64, I32, BOOL, I64]))
    aggs = [AggFunc("Count", [x], distinct=True), AggFunc("Sum", [v]), AggFunc("Sum", [x], distinct=True), AggFunc("Count", [y], distinct=True),
            AggFunc("Count", [b], distinct=True), AggFunc("Max", [v]), AggFunc("Count", [bind_binary_op(x, "*", v)], distinct=True)]
    opts = dict(match_mode=match_mode, count_mode=ffi.COUNT_SQL_ACCUMULATE)
    got, exp = both(lambda l: ex.try_collect(ex.HashAggExecutor(aggs, [k], batches, lib=l, options=l.options(**opts)).execute()), cuda_lib, oracle)
    assert got[0].num_rows == 14  # 13 key values + the NULL key
    assert_batches_match(got, exp)
    got, exp = both(lambda l: ex.try_collect(ex.SimpleAggExecutor(aggs, batches, lib=l, options=l.options(**opts)).execute()), cuda_lib, oracle)
    assert_batches_match(got, exp)
    only_distinct = [AggFunc("Sum", [x], distinct=True), AggFunc("Count", [x], distinct=True)]
    got, exp = both(lambda l: ex.try_collect(ex.SimpleAggExecutor(only_distinct, batches, lib=l, options=l.options(**opts)).execute()), cuda_lib, oracle)
    assert_batches_match(got, exp)
    got, exp = both(lambda l: ex.try_collect(ex.HashAggExecutor([], [k, b], batches, lib=l, options=l.options(**opts)).execute()), cuda_lib, oracle)
    assert_batches_match(got, exp)
    # below a Filter in a plan (fused predicate reaches every sub-operator), and with an ORDER BY on top
    from sqlrs_b200.host.plan import PhysicalFilter, PhysicalHashAgg, PhysicalOrder, PhysicalTableScan

    plan = PhysicalOrder([ex.BoundOrderBy(InputRef(1, I64), asc=False), ex.BoundOrderBy(InputRef(0, I64))],
                         PhysicalHashAgg([AggFunc("Count", [x], distinct=True), AggFunc("Sum", [v])], [k],
                                         PhysicalFilter(bind_binary_op(v, ">", Constant(-3)), PhysicalTableScan(0))))
    table = pa.Table.from_batches(batches).combine_chunks().to_batches()[0]
    for fl in (0, ffi.FLAG_NO_FUSION):
        g, _ = _run_plan(cuda_lib, plan, {0: table.schema}, {0: table}, 2500, flags=fl, **opts)
        e, _ = _run_plan(oracle, plan, {0: table.schema}, {0: table}, 2500, **opts)
        assert_batches_match(g, e)


# ------------------------------------------------------------------ Utf8: dictionary-encoded on ingest, ranks for order
U8 = ffi.DT_UTF8


def _utf8_batch(rng, n, vocab, null_frac=0.1):
    words = [vocab[i] for i in rng.integers(0, len(vocab), n)]
    mask = rng.random(n) < null_frac
    return pa.RecordBatch.from_arrays(
        [pa.array(words, type=pa.string(), mask=mask), pa.array(rng.integers(-50, 50, n).astype(np.int64)),
         pa.array([vocab[i] for i in rng.integers(0, len(vocab), n)], type=pa.string(), mask=rng.random(n) < null_frac)], names=["s", "v", "t"])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_utf8_group_keys_min_max_order_distinct(cuda_lib, oracle, seed):
    """Utf8 group keys (hash_utils.rs:199-208), min_string / max_string (min_max.rs:12-19,47-65), Utf8 sort keys (order.rs:45) and
    DISTINCT over strings.  Later batches bring strings the pool has not seen (the accumulated MIN / MAX are re-ranked), keys include
    the empty string and NULL, strings differ in length / share prefixes / hold multi-byte characters."""
    rng = np.random.default_rng(seed)
    vocab1 = ["", "a", "ab", "abc", "b", "Zebra", "zebra", "CO", "CA", "é", "日本", "a b"]
    vocab2 = vocab1 + ["0", "A", "aa", "zz", "~", "CO ", "ß"]
    batches = [_utf8_batch(rng, 300, vocab1), _utf8_batch(rng, 500, vocab2), _utf8_batch(rng, 7, ["only", ""])]
    s, v, t = InputRef(0, U8), InputRef(1, I64), InputRef(2, U8)
    opts = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)

    def agg(lib):  # group by a Utf8 key: sum / count, and min / max of ANOTHER Utf8 column
        return ex.try_collect(ex.HashAggExecutor([AggFunc("Sum", [v]), AggFunc("Count", [s]), AggFunc("Min", [t]), AggFunc("Max", [t])], [s], batches, lib=lib,
                                                 options=lib.options(**opts)).execute())

    got, exp = both(agg, cuda_lib, oracle)
    assert_batches_match(got, exp)

    def simple(lib):
        return ex.try_collect(ex.SimpleAggExecutor([AggFunc("Max", [s]), AggFunc("Min", [s]), AggFunc("Count", [s])], batches, lib=lib, options=lib.options(**opts)).execute())

    got, exp = both(simple, cuda_lib, oracle)
    assert_batches_match(got, exp)

    def distinct(lib):
        return ex.try_collect(ex.HashAggExecutor([], [s, t], batches, lib=lib, options=lib.options(**opts)).execute())

    got, exp = both(distinct, cuda_lib, oracle)
    assert_batches_match(got, exp)

    for asc in (True, False):
        def order(lib):
            return ex.try_collect(ex.OrderExecutor([ex.BoundOrderBy(s, asc), ex.BoundOrderBy(v, True), ex.BoundOrderBy(t, not asc)], batches, lib=lib).execute())

        got, exp = both(order, cuda_lib, oracle)
        # rows that tie on every key are unpinned in the reference (unstable sort): compare the key columns and the multiset
        assert [(r[0], r[1], r[2]) for r in rows_of(got)] == [(r[0], r[1], r[2]) for r in rows_of(exp)]

    def filt(lib):
        pred = BinaryOp("OR", BinaryOp("=", s, Constant("CO"), BOOL), BinaryOp("<>", t, Constant("never seen before"), BOOL), BOOL)
        return ex.try_collect(ex.FilterExecutor(pred, batches, lib=lib).execute())

    got, exp = both(filt, cuda_lib, oracle)
    assert_batches_match(got, exp)


def test_utf8_ordering_comparisons_in_filters_and_join_filters(cuda_lib, oracle):
    """<, <=, >, >= over Utf8 columns and literals (gt_dyn & co, array_compute.rs:80-83): in a Filter, in a Filter fused below an
    aggregate, and in a non-equi join filter.  A later batch brings strings the pool has not ranked yet (the device rank table
    is rebuilt before the next launch); prefixes, the empty string, multi-byte characters and NULLs are in the mix."""
    rng = np.random.default_rng(11)
    vocab1 = ["", "a", "ab", "abc", "b", "Zebra", "zebra", "CO", "CA", "é", "日本", "a b"]
    vocab2 = vocab1 + ["0", "A", "aa", "zz", "~", "CO ", "ß", "1000", "20"]
    batches = [_utf8_batch(rng, 400, vocab1), _utf8_batch(rng, 600, vocab2)]
    s, v, t = InputRef(0, U8), InputRef(1, I64), InputRef(2, U8)
    BOOL = ffi.DT_BOOL
    for op in ("<", "<=", ">", ">="):
        pred = BinaryOp("OR", BinaryOp(op, s, t, BOOL), BinaryOp("AND", BinaryOp(op, t, Constant("b"), BOOL), bind_binary_op(v, ">", Constant(0)), BOOL), BOOL)
        got, exp = both(lambda l: ex.try_collect(ex.FilterExecutor(pred, batches, lib=l).execute()), cuda_lib, oracle)
        assert_batches_match(got, exp)
        assert 0 < sum(b.num_rows for b in got) < 1000
    # fused below an aggregate (the plan fuses Filter into HashAgg): group by t where s < t
    from sqlrs_b200.host.plan import PhysicalFilter, PhysicalHashAgg, PhysicalTableScan

    schema = batches[0].schema
    plan = PhysicalHashAgg([AggFunc("Sum", [v]), AggFunc("Count", [s])], [t], PhysicalFilter(BinaryOp("<", s, t, BOOL), PhysicalTableScan(0)))
    opts = dict(count_mode=ffi.COUNT_SQL_ACCUMULATE, match_mode=ffi.MATCH_HASH_AND_KEY)
    table = pa.Table.from_batches(batches).combine_chunks().to_batches()[0]
    got, _ = _run_plan(cuda_lib, plan, {0: schema}, {0: table}, 256, **opts)
    exp, _ = _run_plan(oracle, plan, {0: schema}, {0: table}, 256, **opts)
    assert_batches_match(got, exp)
    # non-equi join filter over the joined row: l.s = r.s and l.t < r.t
    left, right = _utf8_batch(rng, 60, vocab1), _utf8_batch(rng, 80, vocab2)
    jschema = pa.schema([pa.field(f"l.{f.name}", f.type) for f in left.schema] + [pa.field(f"r.{f.name}", f.type) for f in right.schema])
    cond = ex.JoinCondition([(InputRef(0, U8), InputRef(0, U8))], filter=BinaryOp("<", InputRef(2, U8), InputRef(5, U8), BOOL))
    for jt in ("Inner", "Left"):
        got, exp = both(lambda l: ex.try_collect(ex.HashJoinExecutor([left], [right], jt, cond, jschema, lib=l, options=l.options(match_mode=ffi.MATCH_HASH_AND_KEY)).execute()),
                        cuda_lib, oracle)
        assert_batches_match(got, exp)


def test_utf8_join_keys_and_payload(cuda_lib, oracle):
    """Utf8 join keys and Utf8 payload columns through every join type (hash_join.rs: keys hashed by create_hashes, payload by take)"""
    rng = np.random.default_rng(5)
    vocab = ["", "x", "xy", "y", "CO", "CA", "Ünï"]
    left, right = _utf8_batch(rng, 40, vocab), _utf8_batch(rng, 60, vocab + ["zz"])
    schema = pa.schema([pa.field(f"l.{f.name}", f.type) for f in left.schema] + [pa.field(f"r.{f.name}", f.type) for f in right.schema])
    cond = ex.JoinCondition([(InputRef(0, U8), InputRef(0, U8))])
    for jt in ("Inner", "Left", "Right", "Full"):
        def join(lib):
            return ex.try_collect(ex.HashJoinExecutor([left], [right], jt, cond, schema, lib=lib, options=lib.options(match_mode=ffi.MATCH_HASH_AND_KEY)).execute())

        got, exp = both(join, cuda_lib, oracle)
        assert_batches_match(got, exp)
