"""The reference's own golden vectors for the hot path, replayed through the C ABI.

Every case runs against the CPU oracle (pins the oracle to the reference; `-m "not gpu"`) and, on the
GPU box, against the CUDA library (`-m gpu`).  Fixtures are the reference's test data:
  hash KAT               src/executor/aggregate/hash_utils.rs:229-247
  HashAgg two chunks     src/executor/aggregate/hash_agg.rs:182-222
  HashJoin 8 tables      src/executor/join/hash_join.rs:393-403,423-750
  executor e2e           src/executor/mod.rs:271-396
  SLT over tests/csv     tests/slt/{aggregation,filter,join,join_filter}.slt
Utf8 columns are outside the CUDA backend's scope: string cases run on the
oracle only, numeric projections of the same tables on both.
"""
import json
import os

import pyarrow as pa
import pytest

from sqlrs_b200.host import executor as ex
from sqlrs_b200.host import ffi
from sqlrs_b200.host.expr import AggFunc, BinaryOp, Constant, InputRef, bind_binary_op
from util import batch, rows_of

I32, I64, F64, BOOL = ffi.DT_INT32, ffi.DT_INT64, ffi.DT_FLOAT64, ffi.DT_BOOL
N = None
# the reference's vectors as a committed fixture (tests/golden/reference_vectors.json, each entry cites its source)
GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.json")))


def _tuples(rows):
    return [tuple(r) for r in rows]


# ---------------------------------------------------------------- fixtures = the reference's data
def employee_numeric():
    """tests/csv/employee.csv, numeric columns: id, salary, department_id (row 4 has NULL salary / department_id)."""
    return batch(["id", "salary", "department_id"], [1, 2, 3, 4], [12000, 10000, 11500, N], [1, 2, 4, N])


def department_numeric():
    return batch(["id"], [1, 2, 3, 4])


def t1():
    return batch(["a", "b", "c"], [0, 1, 2, 2], [4, 5, 7, 8], [7, 8, 9, 1])


def t2():
    return batch(["a", "b", "c"], [10, 20, 30, 40], [2, 2, 3, 4], [7, 5, 6, 6])


def mem_employee():
    """src/executor/mod.rs:220-243 (numeric columns): id, salary"""
    return batch(["id", "salary"], [1, 2, 3, 4], [100, 100, 200, 400], nullable=False)


# ---------------------------------------------------------------- hash KAT
def test_hash_kat(lib):
    """hash_utils.rs:229-247: two identical Float64 columns -> pinned u64 hashes"""
    g = GOLDEN["hash_kat"]
    assert ex.create_hashes([pa.array(c) for c in g["columns"]], lib=lib) == g["hashes"]


def test_hash_single_column_and_null(lib):
    """hash_utils.rs:81-104,167: one column = raw hash_one; a NULL cell keeps the initial 0 (quirk K3)"""
    h = ex.create_hashes([pa.array([0, 1, None, 2, -1, 10471], pa.int64())], lib=lib)
    assert h == [1838465364428186174, 11003429890058878283, 0, 12908271710217070096, 7568711536646566273, 15327478874182264267]


# ---------------------------------------------------------------- HashAgg
def test_hash_agg_two_chunks(lib):
    """hash_agg.rs:182-222: a=[1,1,2], b=[1,1,3] twice; sum(b) group by a -> (1,4),(2,6); header a | Sum(b)"""
    b = batch(["a", "b"], [1, 1, 2], [1, 1, 3], nullable=False)
    op = ex.HashAggExecutor([AggFunc("Sum", [InputRef(1, I64)])], [InputRef(0, I64)], [b, b], lib=lib)
    out = ex.try_collect(op.execute())
    assert len(out) == 1 and out[0].schema.names == ["a", "Sum(b)"]
    assert rows_of(out) == [(1, 4), (2, 6)]


def test_executor_hash_agg(lib):
    """mod.rs:317-350: select salary, count(id), sum(id), max(id), min(id) from employee group by salary"""
    idc, sal = InputRef(0, I64), InputRef(1, I64)
    op = ex.HashAggExecutor([AggFunc("Count", [idc]), AggFunc("Sum", [idc]), AggFunc("Max", [idc]), AggFunc("Min", [idc])], [sal],
                            [mem_employee()], lib=lib)
    out = ex.try_collect(op.execute())
    assert out[0].schema.names == ["salary", "Count(id)", "Sum(id)", "Max(id)", "Min(id)"]
    assert rows_of(out) == [(100, 2, 3, 2, 1), (200, 1, 3, 3, 3), (400, 1, 4, 4, 4)]


def test_executor_simple_agg(lib):
    """mod.rs:293-313: select sum(salary) from employee -> 800"""
    op = ex.SimpleAggExecutor([AggFunc("Sum", [InputRef(1, I64)])], [mem_employee()], lib=lib)
    assert rows_of(ex.try_collect(op.execute())) == [(800,)]


def test_executor_filter(lib):
    """mod.rs:271-291: ... where id = 1 -> the first row"""
    pred = bind_binary_op(InputRef(0, I64), "=", Constant(1))
    out = ex.try_collect(ex.FilterExecutor(pred, [mem_employee()], lib=lib).execute())
    assert rows_of(out) == [(1, 100)]


def test_slt_aggregation_simple(lib):
    """aggregation.slt:1-11 — incl. BASELINE config 1: sum(salary), count(salary) ... where id > 1 -> 21500, 2"""
    emp = employee_numeric()
    idc, sal = InputRef(0, I64), InputRef(1, I64)
    assert rows_of(ex.try_collect(ex.SimpleAggExecutor([AggFunc("Sum", [sal])], [emp], lib=lib).execute())) == [(33500,)]
    filtered = ex.FilterExecutor(bind_binary_op(idc, ">", Constant(1)), [emp], lib=lib).execute()
    aggs = [AggFunc("Sum", [sal]), AggFunc("Sum", [bind_binary_op(idc, "+", Constant(1))]), AggFunc("Count", [idc]), AggFunc("Count", [sal])]
    out = ex.try_collect(ex.SimpleAggExecutor(aggs, filtered, lib=lib).execute())
    assert out[0].schema.names == ["Sum(salary)", "Sum(id+Int64(1))", "Count(id)", "Count(salary)"]
    assert rows_of(out) == [(21500, 12, 3, 2)]
    out = ex.try_collect(ex.SimpleAggExecutor([AggFunc("Max", [sal]), AggFunc("Min", [idc])], [emp], lib=lib).execute())
    assert rows_of(out) == [(12000, 1)]


def test_slt_aggregation_group_by_with_null_group(lib):
    """aggregation.slt:19-26: group by salary — the NULL salary forms its own group, SUM/MAX/MIN of it are NULL"""
    idc, sal = InputRef(0, I64), InputRef(1, I64)
    op = ex.HashAggExecutor([AggFunc("Count", [idc]), AggFunc("Sum", [sal]), AggFunc("Max", [sal]), AggFunc("Min", [sal])], [sal],
                            [employee_numeric()], lib=lib)
    assert rows_of(ex.try_collect(op.execute())) == [(12000, 1, 12000, 12000, 12000), (10000, 1, 10000, 10000, 10000),
                                                      (11500, 1, 11500, 11500, 11500), (N, 1, N, N, N)]


def test_slt_aggregation_two_keys(lib):
    """aggregation.slt:36-43 (numeric projection): group by department_id, id"""
    idc, sal, dep = InputRef(0, I64), InputRef(1, I64), InputRef(2, I64)
    op = ex.HashAggExecutor([AggFunc("Count", [dep]), AggFunc("Sum", [sal])], [dep, idc], [employee_numeric()], lib=lib)
    assert rows_of(ex.try_collect(op.execute())) == [(1, 1, 1, 12000), (2, 2, 1, 10000), (4, 3, 1, 11500), (N, 4, 0, N)]


def test_slt_filter(lib):
    """filter.slt:1-19"""
    emp = employee_numeric()
    idc = InputRef(0, I64)
    gt2 = bind_binary_op(idc, ">", Constant(2))
    ids = lambda pred: [r[0] for r in rows_of(ex.try_collect(ex.FilterExecutor(pred, [emp], lib=lib).execute()))]
    assert ids(gt2) == [3, 4]
    assert ids(BinaryOp("AND", gt2, bind_binary_op(idc, "<", Constant(4)))) == [3]
    assert ids(BinaryOp("OR", bind_binary_op(idc, ">", Constant(3)), bind_binary_op(idc, "=", Constant(1)))) == [1, 4]


# ---------------------------------------------------------------- HashJoin: hash_join.rs tests
def _i32_table(names, *cols):
    return batch(names, *cols, types=[pa.int32()] * 3, nullable=False)


def _join_schema(left, lname, right, rname, join_type):
    lnull, rnull = {"Inner": (False, False), "Left": (False, True), "Right": (True, False), "Full": (True, True)}[join_type]
    fields = [pa.field(f"{lname}.{f.name}", f.type, lnull) for f in left.schema] + [pa.field(f"{rname}.{f.name}", f.type, rnull) for f in right.schema]
    return pa.schema(fields)


JOIN_EXPECTED = {jt: _tuples(GOLDEN["hash_join"][jt]) for jt in ("Inner", "Left", "Right", "Full")}  # hash_join.rs:442-549


@pytest.mark.parametrize("join_type", ["Inner", "Left", "Right", "Full"])
def test_hash_join_results(lib, join_type):
    left = _i32_table(["a1", "b1", "c1"], [0, 1, 2, 3, 4], [0, 4, 5, 5, 8], [10, 7, 8, 9, 10])
    right = _i32_table(["a2", "b1", "c2"], [10, 20, 30], [4, 5, 6], [70, 80, 90])
    schema = _join_schema(left, "l", right, "r", join_type)
    cond = ex.JoinCondition([(InputRef(1, I32), InputRef(1, I32))])
    out = ex.try_collect(ex.HashJoinExecutor([left], [right], join_type, cond, schema, lib=lib).execute())
    assert out[0].schema.names == ["l.a1", "l.b1", "l.c1", "r.a2", "r.b1", "r.c2"]
    assert rows_of(out) == JOIN_EXPECTED[join_type]


JOIN_FILTER_EXPECTED = {jt: _tuples(GOLDEN["hash_join_filter"][jt]) for jt in ("Inner", "Left", "Right", "Full")}  # hash_join.rs:620-748
JOIN_NOFILTER_T1T2 = {  # tests/slt/join_filter.slt: t1.a = t2.b without the non-equi part
    "Inner": [(2, 7, 9, 10, 2, 7), (2, 8, 1, 10, 2, 7), (2, 7, 9, 20, 2, 5), (2, 8, 1, 20, 2, 5)],
    "Left": [(2, 7, 9, 10, 2, 7), (2, 8, 1, 10, 2, 7), (2, 7, 9, 20, 2, 5), (2, 8, 1, 20, 2, 5), (0, 4, 7, N, N, N), (1, 5, 8, N, N, N)],
    "Right": [(2, 7, 9, 10, 2, 7), (2, 8, 1, 10, 2, 7), (2, 7, 9, 20, 2, 5), (2, 8, 1, 20, 2, 5), (N, N, N, 30, 3, 6), (N, N, N, 40, 4, 6)],
    "Full": [(2, 7, 9, 10, 2, 7), (2, 8, 1, 10, 2, 7), (2, 7, 9, 20, 2, 5), (2, 8, 1, 20, 2, 5), (N, N, N, 30, 3, 6), (N, N, N, 40, 4, 6),
             (0, 4, 7, N, N, N), (1, 5, 8, N, N, N)],
}


@pytest.mark.parametrize("join_type", ["Inner", "Left", "Right", "Full"])
def test_hash_join_filter_results_i32(lib, join_type):
    """hash_join.rs:564-748 — Int32 tables, non-equi filter l.c > r.c over the joined row"""
    left = _i32_table(["a", "b", "c"], [0, 1, 2, 2], [4, 5, 7, 8], [7, 8, 9, 1])
    right = _i32_table(["a", "b", "c"], [10, 20, 30, 40], [2, 2, 3, 4], [7, 5, 6, 6])
    schema = _join_schema(left, "l", right, "r", join_type)
    cond = ex.JoinCondition([(InputRef(0, I32), InputRef(1, I32))], BinaryOp(">", InputRef(2, I32), InputRef(5, I32)))
    out = ex.try_collect(ex.HashJoinExecutor([left], [right], join_type, cond, schema, lib=lib).execute())
    assert rows_of(out) == JOIN_FILTER_EXPECTED[join_type]


@pytest.mark.parametrize("join_type", ["Inner", "Left", "Right", "Full"])
@pytest.mark.parametrize("with_filter", [False, True])
def test_slt_join_filter_t1_t2(lib, join_type, with_filter):
    """tests/slt/join_filter.slt:10-91 over tests/csv/t1.csv, t2.csv (Int64 from CSV inference)"""
    left, right = t1(), t2()
    schema = _join_schema(left, "t1", right, "t2", join_type)
    flt = BinaryOp(">", InputRef(2, I64), InputRef(5, I64)) if with_filter else None
    cond = ex.JoinCondition([(InputRef(0, I64), InputRef(1, I64))], flt)
    out = ex.try_collect(ex.HashJoinExecutor([left], [right], join_type, cond, schema, lib=lib).execute())
    assert rows_of(out) == (JOIN_FILTER_EXPECTED if with_filter else JOIN_NOFILTER_T1T2)[join_type]


SLT_EMP_DEPT = {  # tests/slt/join.slt:1-40 projected on (employee.id, employee.department_id, department.id)
    "Left": [(1, 1, 1), (2, 2, 2), (3, 4, 4), (4, N, N)],
    "Right": [(1, 1, 1), (2, 2, 2), (N, N, 3), (3, 4, 4)],
    "Inner": [(1, 1, 1), (2, 2, 2), (3, 4, 4)],
    "Full": [(1, 1, 1), (2, 2, 2), (N, N, 3), (3, 4, 4), (4, N, N)],
}


@pytest.mark.parametrize("join_type", ["Inner", "Left", "Right", "Full"])
def test_slt_join_employee_department(lib, join_type):
    """tests/slt/join.slt:1-40 — a NULL department_id never finds a partner (its hash 0 matches no department)"""
    emp, dep = employee_numeric(), department_numeric()
    schema = _join_schema(emp, "employee", dep, "department", "Full")
    cond = ex.JoinCondition([(InputRef(2, I64), InputRef(0, I64))])
    out = ex.try_collect(ex.HashJoinExecutor([emp], [dep], join_type, cond, schema, lib=lib).execute())
    assert [(r[0], r[2], r[3]) for r in rows_of(out)] == SLT_EMP_DEPT[join_type]


def test_empty_build_side_yields_nothing(lib):
    """hash_join.rs:183-185: no left batch -> empty stream for every join type"""
    right = t2()
    schema = _join_schema(t1(), "t1", right, "t2", "Full")
    cond = ex.JoinCondition([(InputRef(0, I64), InputRef(1, I64))])
    for jt in ("Inner", "Left", "Right", "Full"):
        assert ex.try_collect(ex.HashJoinExecutor([], [right], jt, cond, schema, lib=lib).execute()) == []


# ---------------------------------------------------------------- evaluator
def test_evaluator_input_ref_and_cast(lib):
    """evaluator.rs:93-120"""
    b = batch(["a", "b"], [1, 2, 3], [4, 5, 6], types=[pa.int32(), pa.int32()], nullable=False)
    assert ex.eval_column(InputRef(1, I32), b, lib=lib).to_pylist() == [4, 5, 6]
    from sqlrs_b200.host.expr import TypeCast

    out = ex.eval_column(TypeCast(InputRef(0, I32), I64), b, lib=lib)
    assert out.type == pa.int64() and out.to_pylist() == [1, 2, 3]


# ---------------------------------------------------------------- Utf8 cases of the SLT files: on the CUDA library strings are
# dictionary-encoded on ingest (string pool ids), ordered through the pool's byte-wise ranks, decoded on export
def test_slt_utf8_cases(lib):
    """aggregation.slt:13-17,28-34 (max(last_name); group by state with an empty-string key)"""
    oracle = lib
    emp = pa.RecordBatch.from_arrays(
        [pa.array([1, 2, 3, 4], pa.int64()), pa.array(["Hopkins", "Langford", "Travis", "Mill"]), pa.array(["CA", "CO", "CO", ""]),
         pa.array([12000, 10000, 11500, None], pa.int64())], names=["id", "last_name", "state", "salary"])
    U8 = ffi.DT_UTF8
    out = ex.try_collect(ex.SimpleAggExecutor([AggFunc("Max", [InputRef(3, I64)]), AggFunc("Min", [InputRef(0, I64)]), AggFunc("Max", [InputRef(1, U8)])],
                                              [emp], lib=oracle).execute())
    assert rows_of(out) == [(12000, 1, "Travis")]
    st, sal = InputRef(2, U8), InputRef(3, I64)
    out = ex.try_collect(ex.HashAggExecutor([AggFunc("Count", [st]), AggFunc("Sum", [sal]), AggFunc("Max", [sal]), AggFunc("Min", [sal])], [st], [emp],
                                            lib=oracle).execute())
    assert rows_of(out) == [("CA", 1, 12000, 12000, 12000), ("CO", 2, 21500, 11500, 10000), ("", 1, N, N, N)]


# ---------------------------------------------------------------- Limit / Order / Project
def _range_chunk(lo, hi):
    """limit.rs:121-125 range_to_chunk: one non-nullable Int32 column `a`"""
    return batch(["a"], list(range(lo, hi)), types=[pa.int32()], nullable=False)


@pytest.mark.parametrize("inputs,offset,limit,outputs", [tuple(c) for c in GOLDEN["limit_cases"]["cases"]])  # limit.rs:96-101
def test_limit_executor_cases(lib, inputs, offset, limit, outputs):
    out = ex.try_collect(ex.LimitExecutor(limit, offset, [_range_chunk(*r) for r in inputs], lib=lib).execute())
    assert [b.column(0).to_pylist() for b in out] == [list(range(*r)) for r in outputs]
    for b in out:
        assert b.schema.field(0).name == "a" and b.schema.field(0).type == pa.int32()


def test_executor_limit_and_order(lib):
    """mod.rs:353-396: `select id from employee offset 2 limit 1` -> [3];
    `select id from employee order by id desc offset 2 limit 1` -> [2] (plan: Order -> Project -> Limit, select.rs:34-45)"""
    emp = mem_employee()
    idc = InputRef(0, I64)
    out = ex.try_collect(ex.LimitExecutor(1, 2, ex.ProjectExecutor([idc], [emp], lib=lib).execute(), lib=lib).execute())
    assert len(out) == 1 and out[0].schema.names == ["id"] and rows_of(out) == [(3,)]
    ordered = ex.OrderExecutor([ex.BoundOrderBy(idc, asc=False)], [emp], lib=lib).execute()
    out = ex.try_collect(ex.LimitExecutor(1, 2, ex.ProjectExecutor([idc], ordered, lib=lib).execute(), lib=lib).execute())
    assert len(out) == 1 and rows_of(out) == [(2,)]


def test_slt_limit(lib):
    """limit.slt:1-29 over employee ids 1..4"""
    emp = employee_numeric()
    ids = lambda limit, offset: [r[0] for r in rows_of(ex.try_collect(
        ex.LimitExecutor(limit, offset, ex.ProjectExecutor([InputRef(0, I64)], [emp], lib=lib).execute(), lib=lib).execute()))]
    assert ids(2, 1) == [2, 3]
    assert ids(1, 10) == []
    assert ids(0, 0) == []
    assert ids(None, 2) == [3, 4]
    assert ids(2, None) == [1, 2]


def test_slt_order_numeric(lib):
    """order.slt:1-6: order by id desc offset 2 limit 1 -> 2; two sort keys with directions over numeric columns:
    `order by department_id, id desc` — the NULL department sorts first (SortOptions::default().nulls_first)"""
    emp = employee_numeric()
    idc, dep = InputRef(0, I64), InputRef(2, I64)
    ordered = ex.OrderExecutor([ex.BoundOrderBy(idc, asc=False)], [emp], lib=lib).execute()
    assert rows_of(ex.try_collect(ex.LimitExecutor(1, 2, ex.ProjectExecutor([idc], ordered, lib=lib).execute(), lib=lib).execute())) == [(2,)]
    out = ex.try_collect(ex.OrderExecutor([ex.BoundOrderBy(dep, True), ex.BoundOrderBy(idc, False)], [emp], lib=lib).execute())
    assert len(out) == 1 and out[0].schema.names == ["id", "salary", "department_id"]
    assert [(r[0], r[2]) for r in rows_of(out)] == [(4, N), (1, 1), (2, 2), (3, 4)]
    # descending keeps NULLs first
    out = ex.try_collect(ex.OrderExecutor([ex.BoundOrderBy(dep, False)], [emp], lib=lib).execute())
    assert [(r[0], r[2]) for r in rows_of(out)] == [(4, N), (3, 4), (2, 2), (1, 1)]


def test_order_needs_a_batch(lib):
    """order.rs:27 unwraps the schema of the first batch: an empty child stream is an error"""
    with pytest.raises(ffi.ExecutorError) as err:
        ex.try_collect(ex.OrderExecutor([ex.BoundOrderBy(InputRef(0, I64))], [], lib=lib).execute())
    assert err.value.code == ffi.ERR_INTERNAL


def test_project_field_names(lib):
    """project.rs:20-27 + evaluator.rs:30-64: an InputRef keeps its field, other expressions are named l{op}r / cast names"""
    emp = mem_employee()
    exprs = [InputRef(1, I64), bind_binary_op(InputRef(0, I64), "+", Constant(1)), bind_binary_op(InputRef(1, I64), ">", InputRef(0, I64))]
    out = ex.try_collect(ex.ProjectExecutor(exprs, [emp], lib=lib).execute())
    assert out[0].schema.names == ["salary", "id+Int64(1)", "salary>id"]
    assert not out[0].schema.field(0).nullable and out[0].schema.field(1).nullable
    assert rows_of(out) == [(100, 2, True), (100, 3, True), (200, 4, True), (400, 5, True)]


def test_slt_order_utf8(lib):
    oracle = lib
    """order.slt:8-22: `order by state, id desc` -> 4 (empty) / 1 CA / 3 CO / 2 CO; `order by first_name desc offset 2 limit 1` -> 2"""
    emp = pa.RecordBatch.from_arrays(
        [pa.array([1, 2, 3, 4], pa.int64()), pa.array(["Bill", "Gregg", "John", "Von"]), pa.array(["CA", "CO", "CO", ""])], names=["id", "first_name", "state"])
    U8 = ffi.DT_UTF8
    out = ex.try_collect(ex.OrderExecutor([ex.BoundOrderBy(InputRef(2, U8), True), ex.BoundOrderBy(InputRef(0, I64), False)], [emp], lib=oracle).execute())
    assert [(r[0], r[2]) for r in rows_of(out)] == [(4, ""), (1, "CA"), (3, "CO"), (2, "CO")]
    ordered = ex.OrderExecutor([ex.BoundOrderBy(InputRef(1, U8), False)], [emp], lib=oracle).execute()
    out = ex.try_collect(ex.LimitExecutor(1, 2, ex.ProjectExecutor([InputRef(0, I64)], ordered, lib=oracle).execute(), lib=oracle).execute())
    assert rows_of(out) == [(2,)]


# ---------------------------------------------------------------- DISTINCT, tests/slt/distinct.slt
def test_slt_select_distinct_is_a_group_by_without_aggregates(lib):
    """distinct.slt:9-17: `select distinct a, b from t2` -> HashAgg{group_by: [a, b], agg_funcs: []}, rows in first-appearance order"""
    out = ex.try_collect(ex.HashAggExecutor([], [InputRef(0, I64), InputRef(1, I64)], [t2()], lib=lib).execute())
    assert out[0].schema.names == ["a", "b"]
    assert rows_of(out) == [(10, 2), (20, 2), (30, 3), (40, 4)]


def test_slt_distinct_aggregates(lib):
    """distinct.slt:18-42: sum(distinct b) = 9; sum(distinct b) group by c = 2, 2, 7; count(distinct b) = 3 over t2.b = [2, 2, 3, 4]"""
    b, c = InputRef(1, I64), InputRef(2, I64)
    out = ex.try_collect(ex.SimpleAggExecutor([AggFunc("Sum", [b], distinct=True)], [t2()], lib=lib).execute())
    assert rows_of(out) == [(9,)]
    out = ex.try_collect(ex.HashAggExecutor([AggFunc("Sum", [b], distinct=True)], [c], [t2()], lib=lib).execute())
    assert [r[1] for r in rows_of(out)] == [2, 2, 7] and [r[0] for r in rows_of(out)] == [7, 5, 6]
    out = ex.try_collect(ex.SimpleAggExecutor([AggFunc("Count", [b], distinct=True)], [t2()], lib=lib).execute())
    assert rows_of(out) == [(3,)]


def test_distinct_count_counts_null_as_a_value(lib):
    """count.rs:44-57: DistinctCountAccumulator inserts every ScalarValue, NULL included, into its HashSet — so the NULL
    salary of employee 4 is one more distinct value (unpinned by the reference's tests; the restatement is the definition);
    DistinctSum skips it (sum.rs:125-131 via sum_result), mixed with plain aggregates in one operator"""
    emp = employee_numeric()
    sal, dep = InputRef(1, I64), InputRef(2, I64)
    aggs = [AggFunc("Count", [sal], distinct=True), AggFunc("Count", [sal]), AggFunc("Sum", [sal], distinct=True), AggFunc("Sum", [sal])]
    out = ex.try_collect(ex.SimpleAggExecutor(aggs, [emp, emp], lib=lib, options=lib.options(count_mode=ffi.COUNT_SQL_ACCUMULATE)).execute())
    assert rows_of(out) == [(4, 6, 33500, 67000)]
    out = ex.try_collect(ex.HashAggExecutor([AggFunc("Count", [sal], distinct=True), AggFunc("Sum", [sal], distinct=True), AggFunc("Max", [sal])], [dep],
                                            [emp, emp], lib=lib).execute())
    assert rows_of(out) == [(1, 1, 12000, 12000), (2, 1, 10000, 10000), (4, 1, 11500, 11500), (N, 1, N, N)]


def test_slt_select_distinct_utf8(lib):
    """distinct.slt:1-7: select distinct state from employee -> CA, CO, (empty)"""
    oracle = lib
    emp = pa.RecordBatch.from_arrays([pa.array(["CA", "CO", "CO", ""])], names=["state"])
    out = ex.try_collect(ex.HashAggExecutor([], [InputRef(0, ffi.DT_UTF8)], [emp], lib=oracle).execute())
    assert rows_of(out) == [("CA",), ("CO",), ("",)]


# ---------------------------------------------------------------- resident tables, src/storage/memory.rs
def test_in_memory_storage_resident_tables(lib):
    """memory.rs:38-56,137-170 + its tests :176-214: create_mem_table, get_table, read() returns the batches as given; a plan
    scans the resident table (twice, and two plans) with the results of pushing host batches (mod.rs:293-350)"""
    from sqlrs_b200.host.plan import ExecutorBuilder, PhysicalFilter, PhysicalHashAgg, PhysicalTableScan
    from sqlrs_b200.host.storage import InMemoryStorage, StorageError

    emp = mem_employee()
    storage = InMemoryStorage(lib)
    storage.create_mem_table("employee", [emp.slice(0, 3), emp.slice(3)])
    with pytest.raises(StorageError):
        storage.get_table("nope")
    table = storage.get_table("employee")
    assert table.num_rows == 4 and table.num_batches == 2
    assert rows_of(list(table.read())) == rows_of([emp])
    assert rows_of(list(table.read(projection=[1]))) == [(100,), (100,), (200,), (400,)]
    idc, sal = InputRef(0, I64), InputRef(1, I64)
    plan = PhysicalHashAgg([AggFunc("Count", [idc]), AggFunc("Sum", [idc]), AggFunc("Max", [idc]), AggFunc("Min", [idc])], [sal],
                           PhysicalFilter(bind_binary_op(idc, ">", Constant(0)), PhysicalTableScan(0)))
    opts = lib.options(count_mode=ffi.COUNT_SQL_ACCUMULATE)
    for _ in range(2):  # two plans over the same resident table
        p = ExecutorBuilder(lib, opts).build(plan, {0: emp.schema})
        for _ in range(2):  # and the same plan run twice
            p.push_table_resident(0, table)
            assert rows_of(p.run()) == [(100, 2, 3, 2, 1), (200, 1, 3, 3, 3), (400, 1, 4, 4, 4)]
            p.reset()
        p.close()
    table.close()


# ---------------------------------------------------------------- CrossJoin, src/executor/join/cross_join.rs
def test_slt_cross_join(lib):
    """join.slt:96-103: `select t1.*, t2.* from t1 cross join t2 where t1.a = 0` -> the t1 row (0,4,7) next to every t2 row.
    cross_join.rs:41-55 yields ONE batch per (right batch, left row); an empty left side yields nothing (:33-35)"""
    left = ex.try_collect(ex.FilterExecutor(bind_binary_op(InputRef(0, I64), "=", Constant(0)), [t1()], lib=lib).execute())
    right = t2()
    schema = _join_schema(t1(), "t1", right, "t2", "Inner")
    out = ex.try_collect(ex.CrossJoinExecutor(left, [right], schema, lib=lib).execute())
    assert len(out) == 1 and out[0].schema.names == ["t1.a", "t1.b", "t1.c", "t2.a", "t2.b", "t2.c"]
    assert rows_of(out) == [(0, 4, 7, 10, 2, 7), (0, 4, 7, 20, 2, 5), (0, 4, 7, 30, 3, 6), (0, 4, 7, 40, 4, 6)]
    # all of t1 (two left batches) x t2 in two right batches: 4 left rows x 2 right batches = 8 output batches, left row major
    # within a right batch
    out = ex.try_collect(ex.CrossJoinExecutor([t1().slice(0, 1), t1().slice(1)], [right.slice(0, 3), right.slice(3)], schema, lib=lib).execute())
    assert [b.num_rows for b in out] == [3, 3, 3, 3, 1, 1, 1, 1]
    assert rows_of(out[:4]) == [l + r for l in rows_of([t1()]) for r in rows_of([right.slice(0, 3)])]
    assert rows_of(out[4:]) == [l + (40, 4, 6) for l in rows_of([t1()])]
    assert ex.try_collect(ex.CrossJoinExecutor([], [right], schema, lib=lib).execute()) == []
    # a NULL left cell is repeated as NULL (build_scalar_value_array of a None scalar)
    emp, dep = employee_numeric(), department_numeric()
    es = _join_schema(emp, "employee", dep, "department", "Full")
    out = ex.try_collect(ex.CrossJoinExecutor([emp.slice(3)], [dep], es, lib=lib).execute())
    assert rows_of(out) == [(4, N, N, 1), (4, N, N, 2), (4, N, N, 3), (4, N, N, 4)]


def test_plan_cross_join_then_filter(lib):
    """the plan shape of join.slt:96-103 before predicate push-down: Filter(CrossJoin(scan t1, scan t2))"""
    from sqlrs_b200.host.plan import ExecutorBuilder, PhysicalCrossJoin, PhysicalFilter, PhysicalTableScan

    schema = _join_schema(t1(), "t1", t2(), "t2", "Inner")
    plan = PhysicalFilter(bind_binary_op(InputRef(0, I64), "=", Constant(0)), PhysicalCrossJoin(PhysicalTableScan(0), PhysicalTableScan(1), schema))
    p = ExecutorBuilder(lib).build(plan, {0: t1().schema, 1: t2().schema})
    p.push_table(0, t1())
    p.push_table(1, t2())
    out = p.run()
    p.close()
    assert rows_of(out) == [(0, 4, 7, 10, 2, 7), (0, 4, 7, 20, 2, 5), (0, 4, 7, 30, 3, 6), (0, 4, 7, 40, 4, 6)]


# ---------------------------------------------------------------- v2 engine expression surface
def test_v2_checked_arithmetic_scalar_function_slt(lib):
    """tests/slt/scalar_function.slt:1-37 (a+a, a-a, a*a, a/a over (1),(2),(3),(NULL)) through the *_checked operators of
    src/function/scalar/arithmetic_function.rs; integer overflow is an error there (add_checked & co), not a wrap"""
    from sqlrs_b200.host.expr import fold_and

    b = batch(["a"], [1, 2, 3, None], types=[pa.int32()])
    a = InputRef(0, I32)
    for op, want in (("+checked", [2, 4, 6, N]), ("-checked", [0, 0, 0, N]), ("*checked", [1, 4, 9, N]), ("/checked", [1, 1, 1, N])):
        assert ex.eval_column(BinaryOp(op, a, a, I32), b, lib=lib).to_pylist() == want
    big = batch(["x", "y"], [2**63 - 1, 5, -2**63], [1, 6, -1])
    x, y = InputRef(0, I64), InputRef(1, I64)
    assert ex.eval_column(BinaryOp("+", x, y, I64), big, lib=lib).to_pylist() == [-2**63, 11, 2**63 - 1]  # v1: wrapping (array_compute.rs:76)
    for op in ("+checked", "*checked", "/checked"):
        with pytest.raises(ffi.ExecutorError) as e:
            ex.eval_column(BinaryOp(op, x, y, I64), big, lib=lib)
        assert e.value.code == ffi.ERR_ARROW
    ok = batch(["x", "y"], [2**62, 5, -2**62], [2**62 - 1, 6, -2**62])
    assert ex.eval_column(BinaryOp("+checked", x, y, I64), ok, lib=lib).to_pylist() == [2**63 - 1, 11, -2**63]
    small = batch(["x", "y"], [2**31 - 1, 7], [1, 1], types=[pa.int32(), pa.int32()])
    with pytest.raises(ffi.ExecutorError):
        ex.eval_column(BinaryOp("+checked", InputRef(0, I32), InputRef(1, I32), I32), small, lib=lib)
    # the v2 Filter folds its predicates into one AND conjunction (physical_filter.rs:11-16)
    t = batch(["a", "b"], [1, 2, 3, 4, None], [10, 20, 30, None, 50])
    preds = [bind_binary_op(InputRef(0, I64), ">", Constant(1)), bind_binary_op(InputRef(1, I64), "<", Constant(40))]
    out = ex.try_collect(ex.FilterExecutor(fold_and(preds), [t], lib=lib).execute())
    assert rows_of(out) == [(2, 20), (3, 30)]


def test_comparison_function_slt_strings_compare_bytewise(lib):
    """tests/slt/comparison_function.slt:9-19: `select 100 > 20` = true, `select '1000' > '20'` = false — strings compare byte by
    byte (arrow's gt_utf8 behind gt_dyn, array_compute.rs:80-83), not as numbers; <, <=, >= likewise, NULL propagates"""
    b = batch(["i", "j", "s", "t"], [100, 3], [20, 3], ["1000", "b"], ["20", None], types=[pa.int64(), pa.int64(), pa.string(), pa.string()])
    U8 = ffi.DT_UTF8
    assert ex.eval_column(BinaryOp(">", InputRef(0, I64), InputRef(1, I64), ffi.DT_BOOL), b, lib=lib).to_pylist() == [True, False]
    assert ex.eval_column(BinaryOp(">", InputRef(2, U8), InputRef(3, U8), ffi.DT_BOOL), b, lib=lib).to_pylist() == [False, None]
    assert ex.eval_column(BinaryOp("<", InputRef(2, U8), InputRef(3, U8), ffi.DT_BOOL), b, lib=lib).to_pylist() == [True, None]
    assert ex.eval_column(BinaryOp("<=", InputRef(2, U8), Constant("b"), ffi.DT_BOOL), b, lib=lib).to_pylist() == [True, True]
    assert ex.eval_column(BinaryOp(">=", InputRef(2, U8), Constant("2"), ffi.DT_BOOL), b, lib=lib).to_pylist() == [False, True]


def test_plan_rerun_with_one_table_replaced(lib):
    """sqlrs_plan_clear_table: a plan keeps its pushed tables across runs; one slot is emptied and refilled, the other stays.
    (t1 join t2 on t1.a = t2.b -> nothing matches; t1' join t2 with t1'.a in t2.b -> rows in the reference's order)"""
    from sqlrs_b200.host.plan import ExecutorBuilder, PhysicalHashJoin, PhysicalTableScan

    schema = _join_schema(t1(), "t1", t2(), "t2", "Inner")
    cond = ex.JoinCondition([(InputRef(0, I64), InputRef(1, I64))])
    plan = PhysicalHashJoin(PhysicalTableScan(0), PhysicalTableScan(1), "Inner", cond, schema)
    p = ExecutorBuilder(lib).build(plan, {0: t1().schema, 1: t2().schema})
    p.push_table(0, t1())
    p.push_table(1, t2())
    first = rows_of(p.run())
    assert first == [(2, 7, 9, 10, 2, 7), (2, 8, 1, 10, 2, 7), (2, 7, 9, 20, 2, 5), (2, 8, 1, 20, 2, 5)]
    assert rows_of(p.run()) == first  # nothing is consumed by a run
    p.clear_table(0)
    p.push_table(0, batch(["a", "b", "c"], [4, 3, 9], [1, 1, 1], [5, 6, 7]))
    assert rows_of(p.run()) == [(3, 1, 6, 30, 3, 6), (4, 1, 5, 40, 4, 6)]
    p.close()
