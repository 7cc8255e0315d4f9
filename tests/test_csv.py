"""CSV scan: src/storage/csv.rs:99-235 — header, ',' delimiter, schema inferred from the first 10 records, 1024-row
batches, bounds, projection.  The oracle's CPU reader is pinned by the reference's own fixture (tests/csv/employee.csv through
tests/slt/aggregation.slt:7-11 = BASELINE.json configs[0]) and cross-checked against pyarrow's CSV reader; the CUDA library's
device parser must produce the same batches (`lib` fixture: oracle on CPU, CUDA on the GPU box)."""
import io

import numpy as np
import pyarrow as pa
import pyarrow.csv as pacsv
import pytest

from sqlrs_b200.host import ffi
from sqlrs_b200.host.expr import AggFunc, Constant, InputRef, bind_binary_op
from sqlrs_b200.host.plan import ExecutorBuilder, PhysicalFilter, PhysicalSimpleAgg, PhysicalTableScan
from sqlrs_b200.host.storage import InMemoryStorage
from util import rows_of

# the reference's fixture tests/csv/employee.csv, byte for byte (7 columns; row 4 has an empty state, salary and department_id)
EMPLOYEE_CSV = """id,first_name,last_name,state,job_title,salary,department_id
1,Bill,Hopkins,CA,Manager,12000,1
2,Gregg,Langford,CO,Driver,10000,2
3,John,Travis,CO,"Manager, Software",11500,4
4,Von,Mill,,Defensive End,,
"""
N = None


def test_config1_from_the_csv_file(lib, tmp_path):
    """BASELINE.json configs[0]: SELECT sum(salary), count(salary) FROM employee WHERE id > 1 on tests/csv/employee.csv = 21500, 2
    (aggregation.slt:7-11), CsvScan -> Filter -> SimpleAgg with the table resident after ONE parse"""
    path = tmp_path / "employee.csv"
    path.write_text(EMPLOYEE_CSV)
    storage = InMemoryStorage(lib)
    storage.create_csv_table("employee", str(path))
    table = storage.get_table("employee")
    assert table.num_rows == 4 and table.num_batches == 1
    got = list(table.read())
    assert got[0].schema.names == ["id", "first_name", "last_name", "state", "job_title", "salary", "department_id"]
    assert [str(t) for t in got[0].schema.types] == ["int64", "string", "string", "string", "string", "int64", "int64"]
    assert rows_of(got) == [(1, "Bill", "Hopkins", "CA", "Manager", 12000, 1), (2, "Gregg", "Langford", "CO", "Driver", 10000, 2),
                            (3, "John", "Travis", "CO", "Manager, Software", 11500, 4), (4, "Von", "Mill", "", "Defensive End", N, N)]
    I64 = ffi.DT_INT64
    plan = PhysicalSimpleAgg([AggFunc("Sum", [InputRef(5, I64)]), AggFunc("Count", [InputRef(5, I64)])],
                             PhysicalFilter(bind_binary_op(InputRef(0, I64), ">", Constant(1)), PhysicalTableScan(0)))
    p = ExecutorBuilder(lib, lib.options()).build(plan, {0: got[0].schema})
    p.push_table_resident(0, table)
    assert rows_of(p.run()) == [(21500, 2)]
    p.close()
    # projection + bounds (Table::read(bounds, projection), csv.rs:163-170): offset 1, limit 2 -> rows 2 and 3
    storage.create_csv_table("e2", str(path), projection=[5, 0], bounds=(1, 2))
    assert rows_of(list(storage.get_table("e2").read())) == [(10000, 2), (11500, 3)]


@pytest.mark.parametrize("seed", [0, 1])
def test_csv_random_matches_pyarrow(lib, tmp_path, seed):
    """several thousand rows (multiple 1024-row batches), every inferred type, NULLs, quoted fields with delimiters / quotes / newlines:
    the same values pyarrow's CSV reader produces"""
    rng = np.random.default_rng(seed)
    n = 3000 + seed * 517
    ints = rng.integers(-10**12, 10**12, n)
    floats = np.round(rng.normal(0, 1e4, n), 3)
    words = ["plain", "with,comma", 'say "hi"', "two\nlines", "", "ünï", " spaced "]
    lines = ["i,f,b,s,mixed"]
    rows = []
    for k in range(n):
        i = "" if rng.random() < 0.1 else str(int(ints[k]))
        f = "" if rng.random() < 0.1 else f"{floats[k]:.3f}"
        b = "" if rng.random() < 0.1 else ("true" if rng.random() < 0.5 else "False")
        w = words[int(rng.integers(0, len(words)))]
        s = '"' + w.replace('"', '""') + '"' if any(c in w for c in ',"\n') or w == "" and rng.random() < 0.5 else w
        m = str(int(ints[k]) % 1000) if k % 3 else f"{floats[k]:.3f}"  # ints and decimals: a Float64 column
        lines.append(",".join([i, f, b, s, m]))
        rows.append((N if i == "" else int(i), N if f == "" else float(f), N if b == "" else b.lower() == "true", w, float(m)))
    path = tmp_path / "random.csv"
    path.write_text("\n".join(lines) + "\n")
    storage = InMemoryStorage(lib)
    storage.create_csv_table("t", str(path))
    table = storage.get_table("t")
    got = list(table.read())
    assert table.num_rows == n and [b.num_rows for b in got] == [1024] * (n // 1024) + [n % 1024]
    assert [str(t) for t in got[0].schema.types] == ["int64", "double", "bool", "string", "double"]
    assert rows_of(got) == rows
    ref = pacsv.read_csv(io.BytesIO(path.read_bytes()), convert_options=pacsv.ConvertOptions(strings_can_be_null=False, column_types={"mixed": pa.float64()}),
                         parse_options=pacsv.ParseOptions(newlines_in_values=True))
    assert ref.column("i").to_pylist() == [r[0] for r in rows] and ref.column("f").to_pylist() == [r[1] for r in rows]


def test_csv_missing_file_is_a_storage_error(lib):
    with pytest.raises(ffi.ExecutorError) as e:
        InMemoryStorage(lib).create_csv_table("x", "/nonexistent/file.csv")
    assert e.value.code == ffi.ERR_STORAGE
